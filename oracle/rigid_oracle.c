/*
 * oracle/rigid_oracle.c -- CPU restatement of the articulated rigid-body substep.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under hcr_genesis_lr_cl_b200/ may link, import
 * or call this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg have an independent, sequential statement of the
 * algorithm the CUDA kernel implements (csrc/dynamics_kernel.cuh).
 *
 * PARITY UNPINNED.  In the reference the arithmetic of this stage lives in the
 * third-party engine behind `self._scene.step()`
 * (legged_gym/simulator/genesis_simulator.py:29, genesis-world==0.3.11,
 * pyproject.toml:35), which is neither vendored under /root/reference nor
 * installable here, and the reference ships no golden vectors for it.  This file
 * therefore restates the *published* formulation that engine follows
 * (Featherstone CRBA/RNE in a common world-aligned frame; MuJoCo-style soft
 * constraints: impedance/reference-acceleration rows, regulariser R, projected
 * Gauss-Seidel with friction-cone projection) with the call-site contract of the
 * reference:
 *   - generalized velocity = [base linear (world), base angular (world), joints]
 *     (genesis_simulator.py:130-133,153-157 rely on dofs 0..5 being world-frame),
 *   - applied joint force = PD torque clamped to the URDF effort
 *     (genesis_simulator.py:27),
 *   - get_links_net_contact_force = per-link sum of contact forces, world frame
 *     (genesis_simulator.py:49), get_links_pos/vel = link-frame origin
 *     (genesis_simulator.py:50-51).
 * What is pinned: analytic known-answer tests in tests/test_oracle_physics.py
 * (free fall, CRBA vs finite-difference RNE, momentum conservation, rest force
 * = m g).
 *
 * Build: see oracle/Makefile (REAL=double -> liboracle_f64.so, float -> _f32).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#ifndef REAL
#define REAL double
#endif
typedef REAL real;

#define MAXB 17      /* bodies  (base + chains)            */
#define MAXJ 16      /* actuated joints                    */
#define MAXV 22      /* 6 + MAXJ                            */
#define MAXS 64      /* collision spheres                   */
#define KMAX 8       /* contacts kept per env               */
#define AUXMAX 8     /* joint-limit + frictionloss rows     */
#define RMAX (3 * KMAX + AUXMAX)

/* ---- packed model (layout shared with hcr_genesis_lr_cl_b200/robot_model.py) ---- */
#define BODY_STRIDE 20 /* jp[3] ax[3] com[3] mass I[6](xx yy zz xy xz yz) lo hi effort armature */
typedef struct {
    int C, D, nb, nj, nlinks, nspheres;
    const float *body;      /* [nb][BODY_STRIDE]                      */
    const float *link_off;  /* [nlinks][3]                            */
    const int *link_body;   /* [nlinks]                               */
    const float *sph;       /* [nspheres][4]  pos[3], radius          */
    const int *sph_body;    /* [nspheres]                             */
    const int *sph_link;    /* [nspheres]                             */
} Model;

/* physics parameters, float[16] */
enum { P_DT, P_GRAV, P_TC, P_DAMPRATIO, P_D0, P_DMAX, P_WIDTH, P_MID, P_POWER,
       P_TERRAIN_MU, P_GEOM_MU, P_ITERS, P_HSCALE, P_VSCALE, P_BORDER, P_TOL, P_TRIMESH, P_NPARAM };

static long long g_sweeps, g_substeps, g_rows, g_contacts;
static long long g_hist[64];
static long long g_hist_nc[9][2];   /* by contact count: solves, sweeps */   /* solves by the number of sweeps they took (diagnostics) */
typedef struct { real x, y, z; } v3;
static v3 V(real x, real y, real z) { v3 r = {x, y, z}; return r; }
static v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 scl(v3 a, real s) { return V(a.x * s, a.y * s, a.z * s); }
static real dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
typedef struct { real m[3][3]; } m3;
static v3 mv(const m3 *A, v3 b) {
    return V(A->m[0][0] * b.x + A->m[0][1] * b.y + A->m[0][2] * b.z,
             A->m[1][0] * b.x + A->m[1][1] * b.y + A->m[1][2] * b.z,
             A->m[2][0] * b.x + A->m[2][1] * b.y + A->m[2][2] * b.z);
}
static m3 mm(const m3 *A, const m3 *B) {
    m3 C;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        real s = 0; for (int k = 0; k < 3; k++) s += A->m[i][k] * B->m[k][j]; C.m[i][j] = s; }
    return C;
}
static m3 mt(const m3 *A) { m3 C; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) C.m[i][j] = A->m[j][i]; return C; }
static m3 quat_to_mat(const real q[4]) { /* wxyz */
    real w = q[0], x = q[1], y = q[2], z = q[3]; m3 R;
    R.m[0][0] = 1 - 2 * (y * y + z * z); R.m[0][1] = 2 * (x * y - w * z); R.m[0][2] = 2 * (x * z + w * y);
    R.m[1][0] = 2 * (x * y + w * z); R.m[1][1] = 1 - 2 * (x * x + z * z); R.m[1][2] = 2 * (y * z - w * x);
    R.m[2][0] = 2 * (x * z - w * y); R.m[2][1] = 2 * (y * z + w * x); R.m[2][2] = 1 - 2 * (x * x + y * y);
    return R;
}
static m3 axis_angle(v3 a, real th) { /* Rodrigues, unit axis */
    real c = cos(th), s = sin(th), t = 1 - c; m3 R;
    R.m[0][0] = c + a.x * a.x * t;       R.m[0][1] = a.x * a.y * t - a.z * s; R.m[0][2] = a.x * a.z * t + a.y * s;
    R.m[1][0] = a.y * a.x * t + a.z * s; R.m[1][1] = c + a.y * a.y * t;       R.m[1][2] = a.y * a.z * t - a.x * s;
    R.m[2][0] = a.z * a.x * t - a.y * s; R.m[2][1] = a.z * a.y * t + a.x * s; R.m[2][2] = c + a.z * a.z * t;
    return R;
}

/* spatial inertia about the common reference point O (= base origin, world axes) */
typedef struct { real m; v3 h; m3 I; } sinertia;
static sinertia si_zero(void) { sinertia s; memset(&s, 0, sizeof s); return s; }
static sinertia si_add(sinertia a, const sinertia *b) {
    a.m += b->m; a.h = add(a.h, b->h);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a.I.m[i][j] += b->I.m[i][j];
    return a;
}
/* momentum of motion (w, v): angular L (about O) and linear P */
static void si_apply(const sinertia *s, v3 w, v3 v, v3 *L, v3 *P) {
    *P = add(scl(v, s->m), cross(w, s->h));
    *L = add(mv(&s->I, w), cross(s->h, v));
}

/* ---- kinematic / dynamic quantities of one env ---- */
typedef struct {
    m3 R[MAXB]; v3 o[MAXB];           /* body frames: rotation, origin rel. base origin (world axes) */
    v3 a[MAXB], sv[MAXB];             /* joint motion subspace S = (a, sv = o x a)                      */
    sinertia si[MAXB], comp[MAXB];    /* body / composite inertia about O                              */
    v3 w[MAXB], v[MAXB];              /* spatial velocity (about O)                                     */
} Kin;

static void terrain_query(const int16_t *hf, int rows, int cols, const float *prm, real x, real y, real *h, v3 *n) {
    if (rows == 0) { *h = 0; *n = V(0, 0, 1); return; }
    real hs = prm[P_HSCALE], vs = prm[P_VSCALE];
    real ihs = (real)1 / hs;                       /* grid coordinates by the reciprocal (one division per query, as the kernel) */
    real gx = (x + prm[P_BORDER]) * ihs, gy = (y + prm[P_BORDER]) * ihs;
    int i = (int)floor(gx), j = (int)floor(gy);
    if (i < 0) i = 0; if (i > rows - 2) i = rows - 2;
    if (j < 0) j = 0; if (j > cols - 2) j = cols - 2;
    real u = gx - i, w = gy - j;
    if (u < 0) u = 0; if (u > 1) u = 1; if (w < 0) w = 0; if (w > 1) w = 1;
    real h00 = hf[i * cols + j] * vs, h10 = hf[(i + 1) * cols + j] * vs;
    real h01 = hf[i * cols + j + 1] * vs, h11 = hf[(i + 1) * cols + j + 1] * vs;
    real dhx, dhy;
    if (prm[P_TRIMESH] != 0) {     /* mesh_type "trimesh": cell cut along (i,j)-(i+1,j+1), legged_gym/utils/terrain_utils.py:887-900 */
        if (u >= w) { dhx = h10 - h00; dhy = h11 - h10; } else { dhx = h11 - h01; dhy = h01 - h00; }
        *h = h00 + u * dhx + w * dhy;
    } else if (u + w <= 1) { dhx = h10 - h00; dhy = h01 - h00; *h = h00 + u * dhx + w * dhy; }
    else { dhx = h11 - h01; dhy = h11 - h10; *h = h11 - (1 - u) * dhx - (1 - w) * dhy; }
    v3 g = V(-dhx * ihs, -dhy * ihs, 1);
    *n = scl(g, 1 / sqrt(dot(g, g)));
}

static real impedance(const float *prm, real pos) {
    real x = fabs(pos) / prm[P_WIDTH];
    if (x >= 1) return prm[P_DMAX];
    real mid = prm[P_MID], p = prm[P_POWER], y;
    if (x < mid) y = pow(x, p) / pow(mid, p - 1);
    else y = 1 - pow(1 - x, p) / pow(1 - mid, p - 1);
    return prm[P_D0] + y * (prm[P_DMAX] - prm[P_D0]);
}

static void cholesky(int n, real A[MAXV][MAXV]) { /* in place, lower */
    for (int j = 0; j < n; j++) {
        for (int k = 0; k < j; k++) for (int i = j; i < n; i++) A[i][j] -= A[i][k] * A[j][k];
        real d = sqrt(A[j][j]);
        for (int i = j; i < n; i++) A[i][j] /= d;
    }
}
static void chol_solve(int n, real L[MAXV][MAXV], real *b) {
    for (int i = 0; i < n; i++) { real s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * b[k]; b[i] = s / L[i][i]; }
    for (int i = n - 1; i >= 0; i--) { real s = b[i]; for (int k = i + 1; k < n; k++) s -= L[k][i] * b[k]; b[i] = s / L[i][i]; }
}

static void forward_kinematics(const Model *M, const real quat[4], const real *q, real mass_add, const real com_shift[3], Kin *K) {
    K->R[0] = quat_to_mat(quat); K->o[0] = V(0, 0, 0); K->a[0] = K->sv[0] = V(0, 0, 0);
    for (int b = 0; b < M->nb; b++) {
        const float *B = M->body + b * BODY_STRIDE;
        if (b > 0) {
            int par = ((b - 1) % M->D == 0) ? 0 : b - 1;
            v3 ax = V(B[3], B[4], B[5]);
            K->o[b] = add(K->o[par], mv(&K->R[par], V(B[0], B[1], B[2])));
            K->a[b] = mv(&K->R[par], ax);
            K->sv[b] = cross(K->o[b], K->a[b]);
            m3 Rj = axis_angle(ax, q[b - 1]);
            K->R[b] = mm(&K->R[par], &Rj);
        }
        real m = B[9]; v3 com = V(B[6], B[7], B[8]);
        if (b == 0) { m += mass_add; com = add(com, V(com_shift[0], com_shift[1], com_shift[2])); }
        v3 c = add(K->o[b], mv(&K->R[b], com));
        m3 Il = {{{B[10], B[13], B[14]}, {B[13], B[11], B[15]}, {B[14], B[15], B[12]}}};
        m3 Rt = mt(&K->R[b]); m3 t = mm(&K->R[b], &Il); m3 Iw = mm(&t, &Rt);
        real cc = dot(c, c); real cv[3] = {c.x, c.y, c.z};
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
            Iw.m[i][j] += m * ((i == j ? cc : 0) - cv[i] * cv[j]);
        K->si[b].m = m; K->si[b].h = scl(c, m); K->si[b].I = Iw;
    }
}

/*
 * One physics substep for one env.
 *   st: pos[3] quat_wxyz[4] lin[3] ang[3]  (13)
 *   q, qd: [nj] in model chain order;  tau_cmd: unclamped PD torque
 *   envp: mass_add, com_shift[3], fric_ratio  (5)
 *   jp:  per joint armature[nj], damping[nj], frictionloss[nj]
 * Outputs: link_force[nlinks][3] (world), ncontact.
 */
static void substep_one(const Model *M, const float *prm, const int16_t *hf, int rows, int cols,
                        real *st, real *q, real *qd, const real *tau_cmd, const real *envp, const real *jp,
                        real *link_force, int *ncontact, real *warm /* [48] in/out: PGS warm start, see DESIGN.md */) {
    const int nb = M->nb, nj = M->nj, nv = 6 + nj, D = M->D;
    const real h = prm[P_DT];
    Kin K;
    forward_kinematics(M, st + 3, q, envp[0], envp + 1, &K);
    v3 vb = V(st[7], st[8], st[9]), wb = V(st[10], st[11], st[12]);

    /* --- velocities, bias accelerations, RNE (all about O, world axes) --- */
    v3 aw[MAXB], av[MAXB], fn[MAXB], ff[MAXB];
    K.w[0] = wb; K.v[0] = vb;
    aw[0] = V(0, 0, 0); av[0] = add(cross(vb, wb), V(0, 0, prm[P_GRAV]));
    for (int b = 1; b < nb; b++) {
        int par = ((b - 1) % D == 0) ? 0 : b - 1;
        real qdj = qd[b - 1];
        K.w[b] = add(K.w[par], scl(K.a[b], qdj));
        K.v[b] = add(K.v[par], scl(K.sv[b], qdj));
        /* A += (V x S) qd   (motion cross product) */
        aw[b] = add(aw[par], scl(cross(K.w[b], K.a[b]), qdj));
        av[b] = add(av[par], scl(add(cross(K.w[b], K.sv[b]), cross(K.v[b], K.a[b])), qdj));
    }
    for (int b = 0; b < nb; b++) {
        v3 LA, PA, LV, PV;
        si_apply(&K.si[b], aw[b], av[b], &LA, &PA);
        si_apply(&K.si[b], K.w[b], K.v[b], &LV, &PV);
        fn[b] = add(LA, add(cross(K.w[b], LV), cross(K.v[b], PV)));
        ff[b] = add(PA, cross(K.w[b], PV));
    }
    real bias[MAXV];
    for (int b = nb - 1; b >= 1; b--) {
        int par = ((b - 1) % D == 0) ? 0 : b - 1;
        bias[6 + b - 1] = dot(K.a[b], fn[b]) + dot(K.sv[b], ff[b]);
        fn[par] = add(fn[par], fn[b]); ff[par] = add(ff[par], ff[b]);
    }
    bias[0] = ff[0].x; bias[1] = ff[0].y; bias[2] = ff[0].z;
    bias[3] = fn[0].x; bias[4] = fn[0].y; bias[5] = fn[0].z;

    /* --- CRBA --- */
    real Mm[MAXV][MAXV]; memset(Mm, 0, sizeof Mm);
    for (int b = nb - 1; b >= 0; b--) K.comp[b] = K.si[b];
    for (int b = nb - 1; b >= 1; b--) {
        int par = ((b - 1) % D == 0) ? 0 : b - 1;
        K.comp[par] = si_add(K.comp[par], &K.comp[b]);
    }
    for (int b = 1; b < nb; b++) {
        v3 L, P; si_apply(&K.comp[b], K.a[b], K.sv[b], &L, &P);
        int j = 6 + b - 1;
        Mm[j][j] = dot(K.a[b], L) + dot(K.sv[b], P);
        int i = b;
        while ((i - 1) % D != 0) { i--; Mm[6 + i - 1][j] = Mm[j][6 + i - 1] = dot(K.a[i], L) + dot(K.sv[i], P); }
        real Pv[3] = {P.x, P.y, P.z}, Lv[3] = {L.x, L.y, L.z};
        for (int k = 0; k < 3; k++) { Mm[k][j] = Mm[j][k] = Pv[k]; Mm[3 + k][j] = Mm[j][3 + k] = Lv[k]; }
    }
    {
        const sinertia *c = &K.comp[0];
        for (int k = 0; k < 3; k++) {
            v3 e = V(k == 0, k == 1, k == 2), L, P; real t[3];
            si_apply(c, V(0, 0, 0), e, &L, &P);                 /* linear dof k */
            t[0] = P.x; t[1] = P.y; t[2] = P.z; for (int r = 0; r < 3; r++) Mm[r][k] = t[r];
            t[0] = L.x; t[1] = L.y; t[2] = L.z; for (int r = 0; r < 3; r++) Mm[3 + r][k] = t[r];
            si_apply(c, e, V(0, 0, 0), &L, &P);                 /* angular dof k */
            t[0] = P.x; t[1] = P.y; t[2] = P.z; for (int r = 0; r < 3; r++) Mm[r][3 + k] = t[r];
            t[0] = L.x; t[1] = L.y; t[2] = L.z; for (int r = 0; r < 3; r++) Mm[3 + r][3 + k] = t[r];
        }
    }
    for (int j = 0; j < nj; j++) Mm[6 + j][6 + j] += jp[j] + h * jp[nj + j]; /* armature + implicit damping */

    /* --- smooth acceleration --- */
    real rhs[MAXV], nu[MAXV];
    nu[0] = vb.x; nu[1] = vb.y; nu[2] = vb.z; nu[3] = wb.x; nu[4] = wb.y; nu[5] = wb.z;
    for (int j = 0; j < nj; j++) nu[6 + j] = qd[j];
    for (int k = 0; k < 6; k++) rhs[k] = -bias[k];
    for (int j = 0; j < nj; j++) {
        real eff = M->body[(j + 1) * BODY_STRIDE + 18];
        real t = tau_cmd[j]; if (t > eff) t = eff; if (t < -eff) t = -eff;
        rhs[6 + j] = t - jp[nj + j] * qd[j] - bias[6 + j];
    }
    real Lc[MAXV][MAXV]; memcpy(Lc, Mm, sizeof Mm);
    cholesky(nv, Lc);
    real afree[MAXV]; memcpy(afree, rhs, sizeof rhs); chol_solve(nv, Lc, afree);

    /* --- collision detection: spheres vs terrain --- */
    real sd[MAXS]; v3 sn[MAXS], sx[MAXS]; int act[MAXS], nact = 0;
    for (int s = 0; s < M->nspheres; s++) {
        const float *S = M->sph + 4 * s; int b = M->sph_body[s];
        sx[s] = add(K.o[b], mv(&K.R[b], V(S[0], S[1], S[2])));
        real hh; terrain_query(hf, rows, cols, prm, st[0] + sx[s].x, st[1] + sx[s].y, &hh, &sn[s]);
        sd[s] = (st[2] + sx[s].z - hh) * sn[s].z - S[3];
        act[s] = sd[s] < 0; nact += act[s];
    }
    while (nact > KMAX) { /* drop the shallowest (ties: highest index) */
        int worst = -1;
        for (int s = 0; s < M->nspheres; s++) if (act[s] && (worst < 0 || sd[s] >= sd[worst])) worst = s;
        act[worst] = 0; nact--;
    }

    /* --- constraint rows --- */
    real J[RMAX][MAXV], pos[RMAX], lo[RMAX], hi[RMAX]; int kind[RMAX]; /* 0 normal,1/2 tangent,3 limit,4 frictionloss */
    v3 cdir[RMAX]; int clink[RMAX]; real cmu[KMAX]; int csph[KMAX], auxcode[AUXMAX];
    int R = 0, nc = 0;
    memset(J, 0, sizeof J);
    real mu = prm[P_GEOM_MU] * envp[4]; if (prm[P_TERRAIN_MU] > mu) mu = prm[P_TERRAIN_MU];
    for (int s = 0; s < M->nspheres; s++) if (act[s]) {
        v3 n = sn[s];
        v3 e = (fabs(n.x) < 0.9) ? V(1, 0, 0) : V(0, 1, 0);
        v3 t1 = sub(e, scl(n, dot(e, n))); t1 = scl(t1, 1 / sqrt(dot(t1, t1)));
        v3 t2 = cross(n, t1);
        v3 xc = sub(sx[s], scl(n, M->sph[4 * s + 3] + (real)0.5 * sd[s]));
        v3 dirs[3] = {n, t1, t2};
        int b = M->sph_body[s];
        for (int d = 0; d < 3; d++) {
            v3 dd = dirs[d]; real *Jr = J[R];
            Jr[0] = dd.x; Jr[1] = dd.y; Jr[2] = dd.z;
            v3 xd = cross(xc, dd); Jr[3] = xd.x; Jr[4] = xd.y; Jr[5] = xd.z;
            for (int i = b; i >= 1; i--) { Jr[6 + i - 1] = dot(dd, cross(K.a[i], sub(xc, K.o[i]))); if ((i - 1) % D == 0) break; }
            kind[R] = d; pos[R] = (d == 0) ? sd[s] : 0; cdir[R] = dd; clink[R] = M->sph_link[s];
            lo[R] = 0; hi[R] = 0; R++;
        }
        csph[nc] = s; cmu[nc++] = mu;
    }
    int naux = 0;
    for (int j = 0; j < nj && naux < AUXMAX; j++) {
        real l = M->body[(j + 1) * BODY_STRIDE + 16], u = M->body[(j + 1) * BODY_STRIDE + 17];
        if (q[j] < l) { J[R][6 + j] = 1; pos[R] = q[j] - l; kind[R] = 3; auxcode[naux] = 8 * j + 7; R++; naux++; }
        else if (q[j] > u) { J[R][6 + j] = -1; pos[R] = u - q[j]; kind[R] = 3; auxcode[naux] = 8 * j + 6; R++; naux++; }
    }
    for (int j = 0; j < nj && naux < AUXMAX; j++) if (jp[2 * nj + j] > 0) {
        J[R][6 + j] = 1; pos[R] = 0; kind[R] = 4; lo[R] = -jp[2 * nj + j]; hi[R] = jp[2 * nj + j]; auxcode[naux] = 8 * j + 4; R++; naux++;
    }

    real f[RMAX]; memset(f, 0, sizeof f);
    /* warm start: a contact keeps the force of the same sphere from the previous substep, an aux row that of the same
       (joint, kind, side); anything new starts from zero */
    for (int c = 0; c < nc; c++) for (int k = 0; k < KMAX; k++) if ((int)warm[4 * k] == csph[c] + 1) { f[3 * c] = warm[4 * k + 1]; f[3 * c + 1] = warm[4 * k + 2]; f[3 * c + 2] = warm[4 * k + 3]; }
    for (int a = 0; a < naux; a++) for (int k = 0; k < AUXMAX; k++) if ((int)warm[32 + 2 * k] == auxcode[a] + 1) f[3 * nc + a] = warm[32 + 2 * k + 1];
    real Y[RMAX][MAXV];
    real acc[MAXV]; memcpy(acc, afree, sizeof afree);
    if (R > 0) {
        real A[RMAX][RMAX], bb[RMAX], Rr[RMAX];
        const real tc = prm[P_TC], dr = prm[P_DAMPRATIO], dmax = prm[P_DMAX];
        const real kk = 1 / (dmax * dmax * tc * tc * dr * dr), bd = 2 / (dmax * tc);
        for (int r = 0; r < R; r++) { memcpy(Y[r], J[r], sizeof(real) * MAXV); chol_solve(nv, Lc, Y[r]); }
        for (int r = 0; r < R; r++) for (int s = 0; s < R; s++) { real a = 0; for (int k = 0; k < nv; k++) a += J[r][k] * Y[s][k]; A[r][s] = a; }
        real imp = 0;
        for (int r = 0; r < R; r++) {
            real vel = 0, ja = 0;
            for (int k = 0; k < nv; k++) { vel += J[r][k] * nu[k]; ja += J[r][k] * afree[k]; }
            if (kind[r] == 0 || kind[r] >= 3) imp = impedance(prm, pos[r]); /* tangents reuse their normal's */
            real aref = -bd * vel - kk * imp * pos[r];
            Rr[r] = (1 - imp) / imp * A[r][r];
            bb[r] = ja - aref;
        }
        int iters = (int)prm[P_ITERS];
        real Binv[KMAX][6];      /* per contact: inverse of the 3x3 block A_cc + diag(R_c), upper triangle */
        for (int c = 0; c < nc; c++) {
            const int r = 3 * c;
            const real a00 = A[r][r] + Rr[r], a01 = A[r][r + 1], a02 = A[r][r + 2], a11 = A[r + 1][r + 1] + Rr[r + 1], a12 = A[r + 1][r + 2],
                       a22 = A[r + 2][r + 2] + Rr[r + 2];
            const real c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
            const real idet = 1 / (a00 * c00 + a01 * c01 + a02 * c02);
            Binv[c][0] = c00 * idet; Binv[c][1] = c01 * idet; Binv[c][2] = c02 * idet; Binv[c][3] = (a00 * a22 - a02 * a02) * idet;
            Binv[c][4] = (a01 * a02 - a00 * a12) * idet; Binv[c][5] = (a00 * a11 - a01 * a01) * idet;
        }
        for (int it = 0; it < iters; it++) {
            real fprev[RMAX]; memcpy(fprev, f, sizeof f);
            /* contacts: block Gauss-Seidel, the three rows of a contact solved together against the 3x3 block
               A_cc + diag(R_c), then f_n >= 0 and the tangential pair projected onto the friction disc */
            for (int c = 0; c < nc; c++) {
                const int r = 3 * c;
                real g[3];
                for (int d = 0; d < 3; d++) {
                    g[d] = bb[r + d] + Rr[r + d] * f[r + d];
                    for (int s2 = 0; s2 < R; s2++) g[d] += A[r + d][s2] * f[s2];
                }
                const real *bi = Binv[c];   /* b00 b01 b02 b11 b12 b22 */
                real n0 = f[r] - (bi[0] * g[0] + bi[1] * g[1] + bi[2] * g[2]);
                real n1 = f[r + 1] - (bi[1] * g[0] + bi[3] * g[1] + bi[4] * g[2]);
                real n2 = f[r + 2] - (bi[2] * g[0] + bi[4] * g[1] + bi[5] * g[2]);
                if (n0 < 0) n0 = 0;
                real lim = cmu[c] * n0, t2 = n1 * n1 + n2 * n2;
                if (t2 > lim * lim) { real sc = lim / sqrt(t2); n1 *= sc; n2 *= sc; }
                f[r] = n0; f[r + 1] = n1; f[r + 2] = n2;
            }
            /* joint-limit and frictionloss rows: row-wise */
            for (int r = 3 * nc; r < R; r++) {
                real res = bb[r] + Rr[r] * f[r];
                for (int s = 0; s < R; s++) res += A[r][s] * f[s];
                real fnew = f[r] - res / (A[r][r] + Rr[r]);
                if (kind[r] == 3) { if (fnew < 0) fnew = 0; }
                else if (kind[r] == 4) { if (fnew < lo[r]) fnew = lo[r]; if (fnew > hi[r]) fnew = hi[r]; }
                f[r] = fnew;
            }
            /* convergence: largest change of any row over this sweep, relative to the largest force */
            real dmax = 0, fmax = 0;
            for (int r = 0; r < R; r++) { real d = fabs(f[r] - fprev[r]); if (d > dmax) dmax = d; if (fabs(f[r]) > fmax) fmax = fabs(f[r]); }
            g_sweeps++;
            if (dmax <= prm[P_TOL] * (1 + fmax) || it == iters - 1) { g_hist[it + 1 < 64 ? it + 1 : 63]++; g_hist_nc[nc][0]++; g_hist_nc[nc][1] += it + 1; break; }
        }
        for (int r = 0; r < R; r++) for (int k = 0; k < nv; k++) acc[k] += Y[r][k] * f[r];
    }

    for (int k = 0; k < 48; k++) warm[k] = 0;   /* ids are stored +1 so that 0 means empty */
    for (int c = 0; c < nc; c++) { warm[4 * c] = csph[c] + 1; warm[4 * c + 1] = f[3 * c]; warm[4 * c + 2] = f[3 * c + 1]; warm[4 * c + 3] = f[3 * c + 2]; }
    for (int a = 0; a < naux; a++) { warm[32 + 2 * a] = auxcode[a] + 1; warm[32 + 2 * a + 1] = f[3 * nc + a]; }
    /* --- contact forces per reporting link (world frame) --- */
    for (int i = 0; i < 3 * M->nlinks; i++) link_force[i] = 0;
    for (int r = 0; r < 3 * nc; r++) {
        link_force[3 * clink[r] + 0] += f[r] * cdir[r].x;
        link_force[3 * clink[r] + 1] += f[r] * cdir[r].y;
        link_force[3 * clink[r] + 2] += f[r] * cdir[r].z;
    }
    *ncontact = nc;
    g_substeps++; g_rows += R; g_contacts += nc;

    /* --- semi-implicit Euler --- */
    for (int k = 0; k < nv; k++) nu[k] += h * acc[k];
    for (int k = 0; k < 3; k++) { st[7 + k] = nu[k]; st[10 + k] = nu[3 + k]; st[k] += h * nu[k]; }
    for (int j = 0; j < nj; j++) { qd[j] = nu[6 + j]; q[j] += h * qd[j]; }
    {
        v3 w = V(nu[3], nu[4], nu[5]); real wn = sqrt(dot(w, w)), th = wn * h;
        real dq[4] = {1, 0, 0, 0};
        if (wn > 1e-12) { real s = sin(th / 2) / wn; dq[0] = cos(th / 2); dq[1] = w.x * s; dq[2] = w.y * s; dq[3] = w.z * s; }
        real *Q = st + 3, n[4];
        n[0] = dq[0] * Q[0] - dq[1] * Q[1] - dq[2] * Q[2] - dq[3] * Q[3];
        n[1] = dq[0] * Q[1] + dq[1] * Q[0] + dq[2] * Q[3] - dq[3] * Q[2];
        n[2] = dq[0] * Q[2] - dq[1] * Q[3] + dq[2] * Q[0] + dq[3] * Q[1];
        n[3] = dq[0] * Q[3] + dq[1] * Q[2] - dq[2] * Q[1] + dq[3] * Q[0];
        real nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3]);
        for (int k = 0; k < 4; k++) Q[k] = n[k] / nn;
    }
}

/* link frame-origin positions (world) and linear velocities (world) for the current state */
static void link_kinematics(const Model *M, const real *st, const real *q, const real *qd, real *link_pos, real *link_vel) {
    Kin K; real zero3[3] = {0, 0, 0};
    forward_kinematics(M, st + 3, q, 0, zero3, &K);
    K.w[0] = V(st[10], st[11], st[12]); K.v[0] = V(st[7], st[8], st[9]);
    for (int b = 1; b < M->nb; b++) {
        int par = ((b - 1) % M->D == 0) ? 0 : b - 1;
        K.w[b] = add(K.w[par], scl(K.a[b], qd[b - 1]));
        K.v[b] = add(K.v[par], scl(K.sv[b], qd[b - 1]));
    }
    for (int l = 0; l < M->nlinks; l++) {
        int b = M->link_body[l];
        v3 x = add(K.o[b], mv(&K.R[b], V(M->link_off[3 * l], M->link_off[3 * l + 1], M->link_off[3 * l + 2])));
        v3 vel = add(K.v[b], cross(K.w[b], x));
        link_pos[3 * l] = st[0] + x.x; link_pos[3 * l + 1] = st[1] + x.y; link_pos[3 * l + 2] = st[2] + x.z;
        link_vel[3 * l] = vel.x; link_vel[3 * l + 1] = vel.y; link_vel[3 * l + 2] = vel.z;
    }
}

/* ------------------------------ exported (ctypes) ------------------------------ */
/* diagnostics (single-threaded use): totals since the last oracle_counters() call */
void oracle_sweep_hist_nc(long long *out) { for (int k = 0; k < 9; k++) { out[2 * k] = g_hist_nc[k][0]; out[2 * k + 1] = g_hist_nc[k][1]; g_hist_nc[k][0] = g_hist_nc[k][1] = 0; } }
void oracle_sweep_hist(long long *out) { for (int k = 0; k < 64; k++) { out[k] = g_hist[k]; g_hist[k] = 0; } }
void oracle_counters(long long *out) { out[0] = g_sweeps; out[1] = g_substeps; out[2] = g_rows; out[3] = g_contacts; g_sweeps = g_substeps = g_rows = g_contacts = 0; }

static Model mk_model(const int *mi, const float *mf) {
    Model M; M.C = mi[0]; M.D = mi[1]; M.nb = mi[2]; M.nj = mi[3]; M.nlinks = mi[4]; M.nspheres = mi[5];
    M.link_body = mi + 8; M.sph_body = M.link_body + M.nlinks; M.sph_link = M.sph_body + M.nspheres;
    M.body = mf; M.link_off = M.body + M.nb * BODY_STRIDE; M.sph = M.link_off + 3 * M.nlinks;
    return M;
}

/* terrain height and unit normal under n world points (x, y): the collision surface the substep sees */
void oracle_terrain_query(const float *prm, const int16_t *hf, int rows, int cols, int n, const double *xy, double *h_out, double *n_out) {
    for (int k = 0; k < n; k++) {
        real h; v3 nn;
        terrain_query(hf, rows, cols, prm, (real)xy[2 * k], (real)xy[2 * k + 1], &h, &nn);
        h_out[k] = h; n_out[3 * k] = nn.x; n_out[3 * k + 1] = nn.y; n_out[3 * k + 2] = nn.z;
    }
}

/* All per-env arrays are row-major double [n_envs][k] regardless of REAL (converted on entry/exit). */
void oracle_substep(const int *mi, const float *mf, const float *prm, const int16_t *hf, int rows, int cols, int n_envs,
                    double *state /*[n][13]*/, double *q, double *qd, const double *tau /*[n][nj]*/,
                    const double *envp /*[n][5]*/, const double *jparam /*[n][3*nj]*/,
                    double *link_force /*[n][nlinks][3]*/, int *ncontact /*[n]*/, double *warm /*[n][48]*/) {
    Model M = mk_model(mi, mf); int nj = M.nj;
    for (int e = 0; e < n_envs; e++) {
        real st[13], qq[MAXJ], qv[MAXJ], tt[MAXJ], ep[5], jp[3 * MAXJ], lf[3 * MAXB + 12], wm[48];
        for (int k = 0; k < 13; k++) st[k] = (real)state[e * 13 + k];
        for (int k = 0; k < 48; k++) wm[k] = (real)warm[e * 48 + k];
        for (int k = 0; k < nj; k++) { qq[k] = (real)q[e * nj + k]; qv[k] = (real)qd[e * nj + k]; tt[k] = (real)tau[e * nj + k]; }
        for (int k = 0; k < 5; k++) ep[k] = (real)envp[e * 5 + k];
        for (int k = 0; k < 3 * nj; k++) jp[k] = (real)jparam[e * 3 * nj + k];
        substep_one(&M, prm, hf, rows, cols, st, qq, qv, tt, ep, jp, lf, ncontact + e, wm);
        for (int k = 0; k < 48; k++) warm[e * 48 + k] = wm[k];
        for (int k = 0; k < 13; k++) state[e * 13 + k] = st[k];
        for (int k = 0; k < nj; k++) { q[e * nj + k] = qq[k]; qd[e * nj + k] = qv[k]; }
        for (int k = 0; k < 3 * M.nlinks; k++) link_force[e * 3 * M.nlinks + k] = lf[k];
    }
}

void oracle_link_kinematics(const int *mi, const float *mf, int n_envs, const double *state, const double *q, const double *qd,
                            double *link_pos, double *link_vel) {
    Model M = mk_model(mi, mf); int nj = M.nj;
    for (int e = 0; e < n_envs; e++) {
        real st[13], qq[MAXJ], qv[MAXJ], lp[3 * MAXB + 12], lv[3 * MAXB + 12];
        for (int k = 0; k < 13; k++) st[k] = (real)state[e * 13 + k];
        for (int k = 0; k < nj; k++) { qq[k] = (real)q[e * nj + k]; qv[k] = (real)qd[e * nj + k]; }
        link_kinematics(&M, st, qq, qv, lp, lv);
        for (int k = 0; k < 3 * M.nlinks; k++) { link_pos[e * 3 * M.nlinks + k] = lp[k]; link_vel[e * 3 * M.nlinks + k] = lv[k]; }
    }
}

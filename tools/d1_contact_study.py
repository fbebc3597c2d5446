#!/usr/bin/env python
"""Deviation D1 in numbers: what replacing the URDF's box / cylinder collision geoms by sphere sets, and leaving out
self-collision, costs on the bench workload (ANALYSIS TOOLING, fp64; output committed as profiles/r2_d1_contact_study.txt).

The reference hands Genesis the URDF primitives (Go2: 5 boxes, 17 cylinders, 5 spheres; go2.urdf) with self-collision ON
(genesis_simulator.py:245-255); the backend collides sphere sets (tools/extract_robot_model.py) with the heightfield only.
Genesis itself cannot run here, so the comparison is geometric and kinematic, on states the fp64 oracle visits:

  1. steady-state populations of `go2_ts` envs under N(0,1) actions (the bench workload) and under 0.3 x N(0,1) actions;
  2. for every primitive of every env: the signed distance of the TRUE primitive to the terrain (support point of the box /
     cylinder / sphere along the local terrain normal) next to the signed distance of its sphere set -> contacts the sphere
     set misses, contacts it invents, and the penetration-depth difference where both touch;
  3. per env: primitives in contact vs spheres in contact, how often more than KMAX = 8 spheres are active (contacts the
     solver drops), how many penalised links ("thigh", "calf", ...) touch in either model (the `collision` reward term);
  4. self-collision the backend ignores: pairs of spheres on non-adjacent bodies that overlap (adjacent = same body or
     parent / child, the pairs an engine filters), per env and step.

    python tools/d1_contact_study.py [--envs 512] [--samples 10]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hcr_genesis_lr_cl_b200 import task_spec as T  # noqa: E402
from hcr_genesis_lr_cl_b200.terrain_assets import terrain_for  # noqa: E402
from oracle.cpu_baseline import OracleEnv  # noqa: E402


def quat_to_mat(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rodrigues(a, th):
    """[N] angles about a fixed axis -> [N,3,3]."""
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K[None] + (1 - c) * (K @ K)[None]


def body_frames(model, base_pos, base_quat, q):
    """World rotation [N,nb,3,3] and origin [N,nb,3] of every movable body."""
    N, nb, D = q.shape[0], model.nb, model.chain_len
    R = np.empty((N, nb, 3, 3)); o = np.empty((N, nb, 3))
    R[:, 0] = quat_to_mat(base_quat); o[:, 0] = base_pos
    for b in range(1, nb):
        par = 0 if (b - 1) % D == 0 else b - 1
        B = model.body[b].astype(np.float64)
        o[:, b] = o[:, par] + np.einsum("nij,j->ni", R[:, par], B[0:3])
        R[:, b] = R[:, par] @ rodrigues(B[3:6], q[:, b - 1])
    return R, o


class Terrain:
    def __init__(self, spec, hs):
        self.hs, self.hscale, self.vscale, self.border = hs.astype(np.float64), spec.horizontal_scale, spec.vertical_scale, spec.border_size

    def query(self, x, y):
        """height and unit normal of the cell triangle under (x, y) -- the rule of the dynamics kernel / oracle."""
        gx, gy = (x + self.border) / self.hscale, (y + self.border) / self.hscale
        i = np.clip(np.floor(gx).astype(int), 0, self.hs.shape[0] - 2); j = np.clip(np.floor(gy).astype(int), 0, self.hs.shape[1] - 2)
        u, w = np.clip(gx - i, 0, 1), np.clip(gy - j, 0, 1)
        h00, h01, h10, h11 = (self.hs[i, j] * self.vscale, self.hs[i, j + 1] * self.vscale, self.hs[i + 1, j] * self.vscale, self.hs[i + 1, j + 1] * self.vscale)
        lower = u + w <= 1
        dhx = np.where(lower, h10 - h00, h11 - h01); dhy = np.where(lower, h01 - h00, h11 - h10)
        h = np.where(lower, h00 + u * dhx + w * dhy, h11 - (1 - u) * dhx - (1 - w) * dhy)
        n = np.stack([-dhx / self.hscale, -dhy / self.hscale, np.ones_like(h)], -1)
        return h, n / np.linalg.norm(n, axis=-1, keepdims=True)


def primitive_distance(prim, Rb, ob, terrain):
    """Signed distance [N] of a primitive's deepest point to the terrain (negative = penetrating), normal from the cell
    under the primitive's centre, height re-queried under the support point."""
    c = ob + np.einsum("nij,j->ni", Rb, np.asarray(prim["pos"]))
    Rg = Rb @ np.asarray(prim["rot"])[None]
    _, n = terrain.query(c[:, 0], c[:, 1])
    if prim["type"] == "sphere":
        sup = c - prim["dims"][0] * n
    elif prim["type"] == "box":
        half = np.asarray(prim["dims"]) / 2
        proj = np.einsum("nij,ni->nj", Rg, n)                     # n . axis_k
        sup = c - np.einsum("nij,nj->ni", Rg, np.sign(proj) * half[None])
    else:                                                         # cylinder: axis = local z, dims = (radius, length)
        r, hl = prim["dims"][0], prim["dims"][1] / 2
        a = Rg[:, :, 2]
        na = np.einsum("ni,ni->n", n, a)
        radial = n - na[:, None] * a
        rn = np.linalg.norm(radial, axis=1, keepdims=True)
        radial = np.where(rn > 1e-9, radial / np.maximum(rn, 1e-9), 0.0)
        sup = c - np.sign(na)[:, None] * hl * a - r * radial
    h, n2 = terrain.query(sup[:, 0], sup[:, 1])
    return (sup[:, 2] - h) * n2[:, 2]


def sphere_distances(model_json, R, o, terrain):
    """[N, nspheres] signed distances of the shipped sphere set (the kernel's own rule)."""
    out = []
    for s in model_json["spheres"]:
        c = o[:, s["body"]] + np.einsum("nij,j->ni", R[:, s["body"]], np.asarray(s["pos"]))
        h, n = terrain.query(c[:, 0], c[:, 1])
        out.append((c[:, 2] - h) * n[:, 2] - s["radius"])
    return np.stack(out, 1)


def self_collision_pairs(model_json, model, R, o, neutral=None):
    """[N] number of overlapping sphere pairs on non-adjacent bodies.  Filtered like an engine's broad phase: same body,
    parent / child, and pairs that already overlap in the neutral (default-pose) configuration (`neutral` = their mask)."""
    sp = model_json["spheres"]
    C = np.stack([o[:, s["body"]] + np.einsum("nij,j->ni", R[:, s["body"]], np.asarray(s["pos"])) for s in sp], 1)   # [N,S,3]
    rad = np.array([s["radius"] for s in sp]); body = np.array([s["body"] for s in sp])
    D = model.chain_len
    parent = np.array([-1] + [0 if (b - 1) % D == 0 else b - 1 for b in range(1, model.nb)])
    adj = (body[:, None] == body[None, :]) | (parent[body][:, None] == body[None, :]) | (parent[body][None, :] == body[:, None])
    d = np.linalg.norm(C[:, :, None, :] - C[:, None, :, :], axis=-1) - (rad[:, None] + rad[None, :])[None]
    hit = (d < 0) & ~adj[None] & np.triu(np.ones_like(adj), 1)[None].astype(bool)
    if neutral is not None:
        hit &= ~neutral[None]
    pairs_bodies = set()
    for n_, i, j in zip(*np.nonzero(hit)):
        pairs_bodies.add((model_json["bodies"][body[i]]["name"], model_json["bodies"][body[j]]["name"]))
    return hit.sum((1, 2)), hit, pairs_bodies


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=512)
    ap.add_argument("--samples", type=int, default=10)
    ap.add_argument("--task", default="go2_ts", help="a fused preset: go2_ts (quadruped) or e.g. tron1_pf_ee (biped: the legs can cross)")
    args = ap.parse_args()
    spec = T.PRESETS[args.task]()
    terrain = terrain_for(spec)
    A = spec.num_actions
    mj = json.load(open(os.path.join(ROOT, "hcr_genesis_lr_cl_b200", "assets", ("tron1_pf" if args.task.startswith("tron1") else "go2") + ".json")))
    ter = Terrain(spec, terrain[0])
    links = [l["name"] for l in mj["links"]]
    for scale, label in ((1.0, "N(0,1) actions (bench workload)"), (0.3, "0.3 x N(0,1) actions")):
        env = OracleEnv(spec, args.envs, terrain, precision="f64")
        model = env.model
        pen_links = set(spec.link_groups(model)[1])
        rng = np.random.default_rng(0)
        for _ in range(150):
            env.step(scale * rng.normal(size=(args.envs, A)).astype(np.float32))
        q0 = np.asarray(spec.default_dof_pos, np.float64)[None]
        Rn, on = body_frames(model, np.zeros((1, 3)), np.array([[1.0, 0, 0, 0]]), q0)
        neutral = self_collision_pairs(mj, model, Rn, on)[1][0]                   # pairs overlapping in the default pose: filtered
        acc = {k: [] for k in ("d_true", "d_sph", "type", "n_true", "n_sph", "capped", "pen_true", "pen_sph", "self", "basez")}
        bodies_hit = set()
        for _ in range(args.samples):
            for _ in range(10):
                env.step(scale * rng.normal(size=(args.envs, A)).astype(np.float32))
            st = env.eo.st
            R, o = body_frames(model, st["base_pos"].astype(np.float64), st["base_quat_wxyz"].astype(np.float64), st["q"].astype(np.float64))
            ds = sphere_distances(mj, R, o, ter)
            n_true = np.zeros(args.envs, int); pen_t = np.zeros(args.envs, int); pen_s = np.zeros(args.envs, int)
            seen_t, seen_s = {}, {}
            for prim in mj["primitives"]:
                dt = primitive_distance(prim, R[:, prim["body"]], o[:, prim["body"]], ter)
                idx = [k for k, s in enumerate(mj["spheres"]) if s["src"] == prim["src"] and s["body"] == prim["body"]]
                dsp = ds[:, idx].min(1)
                acc["d_true"].append(dt); acc["d_sph"].append(dsp); acc["type"].append(np.full(args.envs, {"sphere": 0, "box": 1, "cylinder": 2}[prim["type"]]))
                n_true += dt < 0
                seen_t[prim["link"]] = seen_t.get(prim["link"], False) | (dt < 0)
                seen_s[prim["link"]] = seen_s.get(prim["link"], False) | (dsp < 0)
            for l in pen_links:
                pen_t += seen_t.get(l, np.zeros(args.envs, bool)); pen_s += seen_s.get(l, np.zeros(args.envs, bool))
            n_sph = (ds < 0).sum(1)
            sc, _, hitb = self_collision_pairs(mj, model, R, o, neutral)
            bodies_hit |= hitb
            acc["n_true"].append(n_true); acc["n_sph"].append(n_sph); acc["capped"].append(n_sph > 8)
            acc["pen_true"].append(pen_t); acc["pen_sph"].append(pen_s); acc["self"].append(sc); acc["basez"].append(st["base_pos"][:, 2] - st["env_origins"][:, 2])
        dt, dsp, ty = np.concatenate(acc["d_true"]), np.concatenate(acc["d_sph"]), np.concatenate(acc["type"])
        print(f"=== {args.task}, {args.envs} envs x {args.samples} samples, {label}; base height above the env origin {np.concatenate(acc['basez']).mean():.3f} m")
        for t, name in enumerate(("sphere (feet, head)", "box (base, thighs)", "cylinder (hips, calves, head)")):
            m = ty == t
            both, miss, extra = (dt < 0) & (dsp < 0) & m, (dt < 0) & (dsp >= 0) & m, (dt >= 0) & (dsp < 0) & m
            diff = (dsp - dt)[both] * 1e3
            print(f"  {name:30s} touching in both models {both.sum():6d} | only the true primitive {miss.sum():5d} (depth mean {(-dt[miss]).mean() * 1e3 if miss.any() else 0:.2f} mm, "
                  f"max {(-dt[miss]).max() * 1e3 if miss.any() else 0:.2f} mm) | only the sphere set {extra.sum():5d} | "
                  f"depth difference sphere set - primitive: mean {diff.mean() if diff.size else 0:+.2f} mm, p95 |.| {np.quantile(np.abs(diff), 0.95) if diff.size else 0:.2f} mm, max |.| {np.abs(diff).max() if diff.size else 0:.2f} mm")
        nt, ns = np.concatenate(acc["n_true"]), np.concatenate(acc["n_sph"])
        print(f"  contacts per env: true primitives {nt.mean():.2f}, spheres {ns.mean():.2f}; envs with more than 8 active spheres (solver keeps the 8 deepest) {100 * np.concatenate(acc['capped']).mean():.2f} %")
        pt, ps = np.concatenate(acc["pen_true"]), np.concatenate(acc["pen_sph"])
        print(f"  penalised links touching ('collision' reward count): true {pt.mean():.3f}, sphere set {ps.mean():.3f}, envs where the counts differ {100 * (pt != ps).mean():.2f} %")
        sc = np.concatenate(acc["self"])
        print(f"  self-collision (ignored by the backend; same-body, parent/child and neutral-pose pairs [{int(neutral.sum())}] filtered): envs with at least one overlapping sphere pair "
              f"{100 * (sc > 0).mean():.2f} %, pairs per env {sc.mean():.3f}; body pairs seen: {sorted(bodies_hit)[:16]}")


if __name__ == "__main__":
    main()

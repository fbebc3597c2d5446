#!/usr/bin/env python
"""Derive the compact rigid-body model the B200 backend consumes from a robot URDF.

Test/asset tooling (runs in the build container only): reads the URDF that the
reference ships (resources/robots/*/urdf, consumed by the reference at
legged_gym/simulator/genesis_simulator.py:303-311 with merge_fixed_links=True and
links_to_keep) and writes a small JSON with numbers only:

* movable bodies (base + chains of revolute joints) with fixed children merged
  (mass, COM, inertia by the parallel-axis theorem),
* the *reporting* links (movable bodies + links_to_keep, URDF document order; the
  order the reference resolves names against, genesis_simulator.py:333-363),
* collision primitives re-expressed in the owning body's frame and replaced by
  sphere sets (declared deviation, DESIGN.md "contact geometry").

Usage: python tools/extract_robot_model.py <urdf> <out.json> --keep FL_foot ... [--name go2]
"""
import argparse
import json
import xml.etree.ElementTree as ET

import numpy as np


def rpy_to_mat(rpy):
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def vec(s, n=3):
    if s is None:
        return np.zeros(n)
    return np.array([float(x) for x in s.split()])


def origin_of(elem):
    o = elem.find("origin") if elem is not None else None
    if o is None:
        return np.zeros(3), np.eye(3)
    return vec(o.get("xyz")), rpy_to_mat(vec(o.get("rpy")))


def spheres_for_geom(tag, attrib, pos, R, merge_eps=0.012):
    """Replace one primitive by spheres. Returns list of (center[3], radius)."""
    out = []
    if tag == "sphere":
        out.append((pos, float(attrib["radius"])))
    elif tag == "cylinder":
        r, l = float(attrib["radius"]), float(attrib["length"])
        h = max(l / 2 - r, 0.0)
        axis = R[:, 2]
        if h < merge_eps / 2:
            out.append((pos, r))
        else:
            out.append((pos + h * axis, r))
            out.append((pos - h * axis, r))
    elif tag == "box":
        half = vec(attrib["size"]) / 2
        rb = float(half.min())
        ext = half - rb
        cs = []
        for sx in (-1, 1):
            for sy in (-1, 1):
                for sz in (-1, 1):
                    c = np.array([sx, sy, sz]) * ext
                    if not any(np.linalg.norm(c - d) < merge_eps for d in cs):
                        cs.append(c)
        for c in cs:
            out.append((pos + R @ c, rb))
    else:
        raise ValueError(f"unsupported collision primitive {tag}")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("urdf")
    ap.add_argument("out")
    ap.add_argument("--keep", nargs="*", default=[])
    ap.add_argument("--name", default="robot")
    args = ap.parse_args()

    root = ET.parse(args.urdf).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    link_order = [l.get("name") for l in root.findall("link")]
    joints = root.findall("joint")
    parent_joint = {j.find("child").get("link"): j for j in joints}
    base_name = [n for n in link_order if n not in parent_joint][0]

    # movable bodies: base + children of revolute joints, URDF document order
    body_names = [base_name] + [j.find("child").get("link") for j in joints if j.get("type") == "revolute"]
    body_idx = {n: i for i, n in enumerate(body_names)}

    def to_body(name):
        """(body name, p, R): pose of link `name` in the frame of its movable ancestor."""
        p, R = np.zeros(3), np.eye(3)
        while name not in body_idx:
            j = parent_joint[name]
            assert j.get("type") == "fixed", j.get("name")
            jp, jR = origin_of(j)
            p, R = jp + jR @ p, jR @ R
            name = j.find("parent").get("link")
        return name, p, R

    acc = {n: dict(m=0.0, h=np.zeros(3), parts=[]) for n in body_names}
    report, spheres, primitives = [], [], []
    for ln in link_order:
        b, p, R = to_body(ln)
        l = links[ln]
        inr = l.find("inertial")
        if inr is not None and float(inr.find("mass").get("value")) > 0:
            m = float(inr.find("mass").get("value"))
            ip, iR = origin_of(inr)
            I = inr.find("inertia")
            Ic = np.array([[float(I.get("ixx")), float(I.get("ixy")), float(I.get("ixz"))],
                           [float(I.get("ixy")), float(I.get("iyy")), float(I.get("iyz"))],
                           [float(I.get("ixz")), float(I.get("iyz")), float(I.get("izz"))]])
            c = p + R @ ip
            Rc = R @ iR
            acc[b]["m"] += m
            acc[b]["h"] += m * c
            acc[b]["parts"].append((m, c, Rc @ Ic @ Rc.T))
        if ln in body_idx or ln in args.keep:
            report.append(dict(name=ln, body=body_idx[b], offset=p.tolist()))
        for col in l.findall("collision"):
            cp, cR = origin_of(col)
            g = list(col.find("geometry"))[0]
            # the primitive itself, in the owning body's frame (numbers only): what the sphere set stands in for --
            # tools/d1_contact_study.py measures the difference (DESIGN.md deviation D1)
            gp, gR = p + R @ cp, R @ cR
            dims = vec(g.attrib["size"]) if g.tag == "box" else np.array([float(g.attrib["radius"]), float(g.attrib.get("length", 0.0))])
            primitives.append(dict(body=body_idx[b], link_name=(ln if (ln in body_idx or ln in args.keep) else b), type=g.tag,
                                   pos=[float(x) for x in gp], rot=[[float(x) for x in row] for row in gR], dims=[float(x) for x in dims],
                                   src=f"{ln}:{g.tag}"))
            for c, r in spheres_for_geom(g.tag, g.attrib, p + R @ cp, R @ cR):
                spheres.append(dict(body=body_idx[b], link_name=(ln if (ln in body_idx or ln in args.keep) else b),
                                    pos=[float(x) for x in c], radius=float(r), src=f"{ln}:{g.tag}"))

    link_idx = {r["name"]: i for i, r in enumerate(report)}
    for s in spheres + primitives:
        s["link"] = link_idx[s.pop("link_name")]

    bodies = []
    for n in body_names:
        a = acc[n]
        com = a["h"] / a["m"]
        Ic = np.zeros((3, 3))
        for m, c, I in a["parts"]:
            d = c - com
            Ic += I + m * (d @ d * np.eye(3) - np.outer(d, d))
        b = dict(name=n, mass=a["m"], com=com.tolist(),
                 inertia=[Ic[0, 0], Ic[1, 1], Ic[2, 2], Ic[0, 1], Ic[0, 2], Ic[1, 2]])
        if n == base_name:
            b.update(parent=-1, joint_name=None, joint_pos=[0, 0, 0], joint_axis=[0, 0, 0],
                     limit=[0, 0], effort=0.0, velocity=0.0)
        else:
            j = parent_joint[n]
            jp, jR = origin_of(j)
            assert np.allclose(jR, np.eye(3)), "revolute joint origins with rpy are not supported"
            pb, pp, pR = to_body(j.find("parent").get("link"))
            lim = j.find("limit")
            b.update(parent=body_idx[pb], joint_name=j.get("name"), joint_pos=(pp + pR @ jp).tolist(),
                     joint_axis=(pR @ vec(j.find("axis").get("xyz"))).tolist(),
                     limit=[float(lim.get("lower")), float(lim.get("upper"))],
                     effort=float(lim.get("effort")), velocity=float(lim.get("velocity")))
        bodies.append(b)

    # chains: every non-base body must be on a serial chain hanging off the base
    chains = []
    for i, b in enumerate(bodies):
        if b["parent"] == 0:
            chains.append([i])
    for ch in chains:
        while True:
            kids = [i for i, b in enumerate(bodies) if b["parent"] == ch[-1]]
            if not kids:
                break
            assert len(kids) == 1, "branching below the base is not supported"
            ch.append(kids[0])
    assert len({len(c) for c in chains}) == 1
    model = dict(name=args.name, source=args.urdf.split("resources/")[-1], base_link=base_name,
                 num_chains=len(chains), chain_len=len(chains[0]), chains=chains,
                 bodies=bodies, links=report, spheres=spheres, primitives=primitives,
                 total_mass=sum(b["mass"] for b in bodies))
    with open(args.out, "w") as f:
        json.dump(model, f, indent=1)
    print(f"{args.name}: {len(bodies)} bodies, {len(report)} links, {len(spheres)} spheres, mass {model['total_mass']:.4f}")


if __name__ == "__main__":
    main()

"""Generate invesalius3_b200/csrc/mc_tables.h — the 256-case marching-cubes table.

The table is derived here from first principles (no third-party table is copied):

  corners  k = cx | cy<<1 | cz<<2, position (cx, cy, cz) relative to the cell origin
  edges    e = 4*a + (cu | cv<<1): axis a (0=x,1=y,2=z), (cu, cv) = the cell-local
           coordinates on the two other axes in increasing axis order. The edge runs from
           the corner with coordinate 0 on axis a to the one with coordinate 1, and is
           OWNED by voxel origin + (cu, cv placed on their axes), axis a.
  case     bit k set  <=>  corner k is inside (scalar >= iso)

For every case: on each of the 6 faces the crossing edges are paired into segments (two
crossings: joined; four crossings = ambiguous face: each INSIDE corner is cut off on its
own, a rule that depends only on the face's four corner states, hence both cells sharing
the face agree and the surface is watertight). Segments are directed with the inside on
the right when the face is seen from outside the cell; chaining them gives closed
oriented loops, each triangulated as a fan whose apex is the first loop vertex (from the
smallest edge id on) that yields no diagonal inside a cell face. Triangle normals
(right-hand rule, index space x,y,z) point from inside (>= iso) to outside.
"""
from __future__ import annotations

import itertools
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parents[1] / "invesalius3_b200" / "csrc" / "mc_tables.h"


def corner_pos(k):
    return np.array([k & 1, (k >> 1) & 1, (k >> 2) & 1], dtype=float)


def edge_def(e):
    a, j = divmod(e, 4)
    u, v = [ax for ax in range(3) if ax != a]
    base = np.zeros(3)
    base[u] = j & 1
    base[v] = (j >> 1) & 1
    end = base.copy()
    end[a] = 1
    return base, end


def corner_id(p):
    return int(p[0]) | int(p[1]) << 1 | int(p[2]) << 2


EDGE_CORNERS = [(corner_id(edge_def(e)[0]), corner_id(edge_def(e)[1])) for e in range(12)]
EDGE_MID = [(edge_def(e)[0] + edge_def(e)[1]) / 2 for e in range(12)]


def faces():
    out = []
    for ax in range(3):
        for side in (0, 1):
            n = np.zeros(3)
            n[ax] = 1 if side else -1
            corners = [k for k in range(8) if corner_pos(k)[ax] == side]
            edges = [e for e in range(12) if all(corner_pos(c)[ax] == side for c in EDGE_CORNERS[e])]
            out.append((n, corners, edges))
    return out


FACES = faces()


def segments_for_case(case):
    inside = [(case >> k) & 1 for k in range(8)]
    crossing = [inside[a] != inside[b] for a, b in EDGE_CORNERS]
    segs = []  # directed (from_edge, to_edge)
    for n, corners, edges in FACES:
        ce = [e for e in edges if crossing[e]]
        if not ce:
            continue
        pairs = []
        if len(ce) == 2:
            cin = [k for k in corners if inside[k]]
            pairs.append((ce[0], ce[1], np.mean([corner_pos(k) for k in cin], axis=0)))
        else:
            assert len(ce) == 4
            for k in corners:
                if inside[k]:  # cut off each inside corner: the two face edges touching it
                    es = [e for e in ce if k in EDGE_CORNERS[e]]
                    assert len(es) == 2
                    pairs.append((es[0], es[1], corner_pos(k)))
        for ea, eb, pin in pairs:
            d = EDGE_MID[eb] - EDGE_MID[ea]
            mid = (EDGE_MID[ea] + EDGE_MID[eb]) / 2
            left = np.cross(n, d)
            if np.dot(pin - mid, left) > 0:   # inside is on the left -> reverse
                ea, eb = eb, ea
            segs.append((ea, eb))
    return segs


def loops_for_case(case):
    segs = segments_for_case(case)
    nxt = {}
    for a, b in segs:
        assert a not in nxt, (case, segs)
        nxt[a] = b
    assert sorted(nxt) == sorted(nxt.values())
    loops, seen = [], set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop, e = [], start
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == start and len(loop) >= 3
        loops.append(loop)
    return loops


FACE_EDGE_SETS = [set(e) for _, _, e in FACES]


def same_face(a, b):
    return any(a in f and b in f for f in FACE_EDGE_SETS)


def pick_apex(loop):
    """Rotate the loop to the first apex (starting at the smallest edge id) whose fan has no
    diagonal lying inside a cell face: such a diagonal could coincide with the neighbouring
    cell's and make the shared edge non-manifold."""
    n = len(loop)
    for s in range(n):
        rot = loop[s:] + loop[:s]
        if all(not same_face(rot[0], rot[i]) for i in range(2, n - 1)):
            return rot
    raise AssertionError(f"no face-diagonal-free fan for {loop}")


def build():
    tris = []
    for case in range(256):
        t = []
        for loop in loops_for_case(case):
            loop = pick_apex(loop)
            for i in range(1, len(loop) - 1):
                t.append((loop[0], loop[i], loop[i + 1]))
        tris.append(t)
    return tris


def check_orientation(tris):
    """Every triangle's normal must point away from the inside corners it is nearest to:
    checked on the 8 single-corner cases and their complements analytically."""
    for k in range(8):
        for case, sign in ((1 << k, 1.0), (255 ^ (1 << k), -1.0)):
            (a, b, c), = tris[case]
            pa, pb, pc = EDGE_MID[a], EDGE_MID[b], EDGE_MID[c]
            nrm = np.cross(pb - pa, pc - pa)
            cen = (pa + pb + pc) / 3
            assert sign * np.dot(nrm, cen - corner_pos(k)) > 0, (case, k)


def main():
    tris = build()
    check_orientation(tris)
    maxt = max(len(t) for t in tris)
    lines = [
        "// GENERATED by tools/gen_mc_tables.py — do not edit.",
        "// 256-case marching-cubes triangulation derived from first principles (see the generator",
        "// for the corner / edge numbering and the ambiguous-face rule).",
        "#pragma once",
        "#ifndef B2V_MC_QUAL  // mc.cu defines this as __device__",
        "#define B2V_MC_QUAL",
        "#endif",
        f"#define B2V_MC_MAXTRI {maxt}",
        "// number of triangles per case",
        "B2V_MC_QUAL static const unsigned char B2V_MC_NTRI[256] = {",
    ]
    for i in range(0, 256, 32):
        lines.append("  " + ", ".join(str(len(t)) for t in tris[i:i + 32]) + ",")
    lines.append("};")
    lines.append("// edge ids (0..11) of each triangle's corners, -1 padded")
    lines.append(f"B2V_MC_QUAL static const signed char B2V_MC_TRI[256][{3 * maxt}] = {{")
    for case, t in enumerate(tris):
        flat = [e for tri in t for e in tri]
        flat += [-1] * (3 * maxt - len(flat))
        lines.append("  {" + ", ".join(f"{e:2d}" for e in flat) + "},")
    lines.append("};")
    # bit e set <=> edge e is crossed in this case
    masks = []
    for case in range(256):
        inside = [(case >> k) & 1 for k in range(8)]
        masks.append(sum(1 << e for e, (a, b) in enumerate(EDGE_CORNERS) if inside[a] != inside[b]))
    lines.append("// bit e set <=> edge e crosses the iso-surface")
    lines.append("B2V_MC_QUAL static const unsigned short B2V_MC_EDGEMASK[256] = {")
    for i in range(0, 256, 16):
        lines.append("  " + ", ".join(f"0x{m:03x}" for m in masks[i:i + 16]) + ",")
    lines.append("};")
    OUT.write_text("\n".join(lines) + "\n")
    hist = np.bincount([len(t) for t in tris])
    print(f"wrote {OUT}; max triangles per case {maxt}; histogram {hist.tolist()}")


if __name__ == "__main__":
    main()

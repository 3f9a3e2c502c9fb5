"""Generate tests/golden/exact_alpha_64.npz from the reference's golden fields.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py

Source: tutorials/test/exactSolutions/{0,0.25,...,3}/alpha.water.exact -- 13 OpenFOAM
binary volScalarFields (arch "LSB;label=32;scalar=64", 262144 doubles) holding the
"exact" volume fractions of the 3-D deformation test on the 64^3 mesh AFTER
`renumberMesh` (tutorials/test/plicVofAdvectionFoam/Allrun:10-12).  They are mapped
back to blockMesh's natural cell order c = i + 64 j + 4096 k by re-stating OpenFOAM's
default Cuthill-McKee renumbering (bandCompression) and
stored sparsely (cells that are neither 0 nor 1 + the list of full cells).

The mapping is validated here against the reference's own exact sphere/hex overlap
library (oracle/_ref): golden t=0 must agree with sphere(0.35,0.35,0.35; r=0.15) in
every cell to the icosphere-vs-sphere difference (max 1.6e-4).
"""
import os
import re
import sys
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference/tutorials/test/exactSolutions"
TIMES = ["0", "0.25", "0.5", "0.75", "1", "1.25", "1.5", "1.75", "2", "2.25", "2.5", "2.75", "3"]
N = 64


def read_foam_binary_scalar_field(path):
    raw = open(path, "rb").read()
    m = re.search(rb"internalField\s+nonuniform\s+List<scalar>\s*\n?(\d+)\s*\n?\(", raw)
    n = int(m.group(1))
    start = m.end()
    return np.frombuffer(raw, dtype="<f8", count=n, offset=start).copy()


def cuthill_mckee_hex(n):
    """OpenFOAM bandCompression on the n^3 hex mesh in natural order: newToOld[]."""
    nc = n ** 3
    idx = np.arange(nc)
    i, j, k = idx % n, (idx // n) % n, idx // (n * n)
    nbrs = [[] for _ in range(nc)]
    # cellCells order of natural cell c: [c-n^2, c-n, c-1, c+1, c+n, c+n^2]
    cand = [(k > 0, -n * n), (j > 0, -n), (i > 0, -1), (i < n - 1, 1), (j < n - 1, n), (k < n - 1, n * n)]
    deg = np.zeros(nc, dtype=np.int64)
    for ok, d in cand:
        deg += ok
    for c in range(nc):
        l = nbrs[c]
        for ok, d in cand:
            if ok[c]:
                l.append(c + d)
    visited = np.zeros(nc, dtype=bool)
    new_to_old = []
    # start from the lowest-index cell of minimum neighbour count
    while len(new_to_old) < nc:
        unv = np.nonzero(~visited)[0]
        start = unv[np.argmin(deg[unv])]
        q = deque([start])
        while q:
            c = q.popleft()
            if visited[c]:
                continue
            visited[c] = True
            new_to_old.append(c)
            nb = [x for x in nbrs[c] if not visited[x]]
            nb.sort(key=lambda x: deg[x])  # stable ascending neighbour count
            q.extend(nb)
    return np.array(new_to_old, dtype=np.int64)


def main():
    from common import exact_sphere_alpha, meshmod
    new_to_old = cuthill_mckee_hex(N)
    out = {"times": np.array([float(t) for t in TIMES]), "n": np.array(N)}
    sphere = exact_sphere_alpha(meshmod.hex_block(N))
    for t in TIMES:
        g = read_foam_binary_scalar_field(os.path.join(REF, t, "alpha.water.exact"))
        nat = np.empty_like(g)
        nat[new_to_old] = g          # golden[new] belongs to natural cell newToOld[new]
        if t == "0":
            d = np.abs(nat - sphere).max()
            print("t=0: max |golden - exact sphere| after un-renumbering = %.3e" % d)
            assert d < 2e-4, "Cuthill-McKee restatement does not reproduce the golden ordering"
        part = np.nonzero((nat != 0.0) & (nat != 1.0))[0].astype(np.int32)
        full = np.nonzero(nat == 1.0)[0].astype(np.int32)
        out["part_idx_" + t] = part
        out["part_val_" + t] = nat[part]
        out["full_idx_" + t] = full
        mixed = ((nat > 1e-8) & (nat < 1 - 1e-8)).sum()
        print("t=%s: sum(alpha V)=%.13g mixed=%d full=%d" % (t, nat.sum() / N ** 3, mixed, full.size))
    np.savez_compressed(os.path.join(HERE, "exact_alpha_64.npz"), **out)
    print("wrote", os.path.join(HERE, "exact_alpha_64.npz"), os.path.getsize(os.path.join(HERE, "exact_alpha_64.npz")), "bytes")


if __name__ == "__main__":
    main()

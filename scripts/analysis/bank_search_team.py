"""Padding search for the team (plane + zeta-pencil) kernel: plane lanes (s, X, k) read plane k of tile (X*NEQ+e)
and write plane k of B tile X; both must stay <= 2 wavefronts per warp-wide 64-bit access."""
import sys
N, EPB, NEQ = int(sys.argv[1]), int(sys.argv[2]), 5
NP = N ** 3
def mult(vals):
    c = {}
    for v in vals: c[v % 16] = c.get(v % 16, 0) + 1
    return max(c.values())
for ES in range(NP, NP + 17):
    for GB in range(EPB * ES, EPB * ES + 17):
        ld = mult([(X * NEQ) * GB + s * ES + N * N * k for s in range(EPB) for X in range(3) for k in range(N)])
        st = mult([X * GB + s * ES + N * N * k for s in range(EPB) for X in range(3) for k in range(N)])
        if ld <= 2 and st <= 2: print("ES", ES, "GB", GB, "ld", ld, "st", st)

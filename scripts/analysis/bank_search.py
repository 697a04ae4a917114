"""Shared-memory bank-conflict search for the group-pencil element kernel (64-bit accesses: 16 bank pairs;
a warp-wide access needs max-multiplicity-over-bank-pairs wavefronts, 2 is the floor for 32 lanes)."""
import itertools, sys
def wavefronts(addrs):
    cnt = {}
    for a in set(addrs):
        cnt[a % 16] = cnt.get(a % 16, 0) + 1
    return max(cnt.values()) if cnt else 0
def evaluate(N, EPB, PJ, PK, ES, verbose=False):
    NC = N * N
    NT = (EPB * NC + 31) // 32 * 32
    tot = {"xi": 0, "eta": 0, "zeta": 0, "flux": 0}
    ideal = 0
    for w in range(NT // 32):
        lanes = [p for p in range(32 * w, 32 * w + 32) if p < EPB * NC]
        if not lanes: continue
        ideal += 2 if len(lanes) > 16 else 1
        for name in ("xi", "eta", "zeta"):
            ad = []
            for p in lanes:
                s, c = divmod(p, NC); c0, c1 = c % N, c // N
                if name == "xi": b = s * ES + PJ * c0 + PK * c1
                elif name == "eta": b = s * ES + c0 + PK * c1
                else: b = s * ES + c0 + PJ * c1
                ad.append(b)
            tot[name] += wavefronts(ad)
    # flux-phase stores: node-parallel, node n = r*NT + t of the group, n -> (s, l)
    NP = N ** 3
    nn = EPB * NP
    R = (nn + NT - 1) // NT
    fl_ideal = 0
    for r in range(R):
        for w in range(NT // 32):
            ad = []
            for t in range(32 * w, 32 * w + 32):
                n = r * NT + t
                if n >= nn: continue
                s, l = divmod(n, NP)
                i, j, k = l % N, (l // N) % N, l // (N * N)
                ad.append(s * ES + i + PJ * j + PK * k)
            if ad:
                tot["flux"] += wavefronts(ad)
                fl_ideal += 2 if len(ad) > 16 else 1
    return tot, ideal, fl_ideal
if __name__ == "__main__":
    N, EPB = int(sys.argv[1]), int(sys.argv[2])
    best = []
    for PJ in range(N, N + 3):
        for PK in range(PJ * (N - 1) + N, PJ * (N - 1) + N + 20):
            FSmin = PK * (N - 1) + PJ * (N - 1) + N
            for ES in range(FSmin, FSmin + 17):
                tot, ideal, fl = evaluate(N, EPB, PJ, PK, ES)
                score = tot["xi"] + tot["eta"] + tot["zeta"]
                best.append((score, tot["flux"], PJ, PK, ES, tot, ideal, fl))
    best.sort(key=lambda x: (x[0], x[4], x[1]))
    for b in best[:12]: print(b)

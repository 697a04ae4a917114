"""Per-pass digit-order search: the flattened pencil id p of a CTA decodes per pass into (s, c0, c1) with any digit order."""
import itertools, sys
def wavefronts(addrs):
    cnt = {}
    for a in set(addrs):
        cnt[a % 16] = cnt.get(a % 16, 0) + 1
    return max(cnt.values()) if cnt else 0
def decode(p, order, N, EPB):
    # order: tuple of digit names fastest->slowest among 's','a','b'
    rad = {'s': EPB, 'a': N, 'b': N}
    v = {}
    for d in order:
        v[d] = p % rad[d]; p //= rad[d]
    return v['s'], v['a'], v['b']
def pass_cost(N, EPB, name, order, PJ, PK, ES):
    NC = N * N; NT = (EPB * NC + 31) // 32 * 32
    tot = 0
    for w in range(NT // 32):
        ad = []
        for p in range(32 * w, 32 * w + 32):
            if p >= EPB * NC: continue
            s, c0, c1 = decode(p, order, N, EPB)
            if name == "xi": b = s * ES + PJ * c0 + PK * c1
            elif name == "eta": b = s * ES + c0 + PK * c1
            else: b = s * ES + c0 + PJ * c1
            ad.append(b)
        tot += wavefronts(ad)
    return tot
N, EPB = int(sys.argv[1]), int(sys.argv[2])
orders = list(itertools.permutations(('s', 'a', 'b')))
res = []
for PJ in range(N, N + 5):
    for PK in range(PJ * (N - 1) + N, PJ * (N - 1) + N + 24):
        FSmin = PK * (N - 1) + PJ * (N - 1) + N
        for ES in range(FSmin, FSmin + 17):
            sc = 0; pick = {}
            for name in ("xi", "eta", "zeta"):
                c, o = min((pass_cost(N, EPB, name, o, PJ, PK, ES), o) for o in orders)
                sc += c; pick[name] = (c, ''.join(o))
            res.append((sc, ES, PJ, PK, pick))
res.sort(key=lambda x: (x[0], x[1]))
for r in res[:10]: print(r)

"""Aggregate a NRW_GEMM_TIMING_DUMP csv: per GEMM configuration count, time, MMA rate and HBM rate."""
import sys, collections
rows = collections.OrderedDict()
for l in open(sys.argv[1]):
    M, N, K, P, mn, ks, epi, by, ms = l.strip().split(",")
    k = (int(M), int(N), int(K), int(P), int(mn), int(ks), int(epi))
    r = rows.setdefault(k, [0, 0.0, float(by)])
    r[0] += 1; r[1] += float(ms)
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot = sum(r[1] for r in rows.values())
print(f"total {tot/steps:.2f} ms/step in {sum(r[0] for r in rows.values())/steps:.0f} launches")
print("     M     N     K P mn ks  epi |   n/step  ms/step   us/launch  MMA TF/s  GB/s(alg)")
for k, r in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    M, N, K, P, mn, ks, epi = k
    us = r[1] / r[0] * 1e3
    npr = {1: 1, 2: 3, 3: 6}[P]
    print(f"{M:7d} {N:5d} {K:6d} {P} {mn:2d} {ks:3d} {epi:4d} | {r[0]/steps:7.1f} {r[1]/steps:8.2f} {us:10.1f} {2.0*M*N*K*npr/us/1e6:9.0f} {r[2]/us/1e3:9.0f}")

"""cProfile of the reference call sequence through the drop-in shim (see shim_latency.py): where the host time goes."""
import cProfile, io, pstats, sys
sys.path.insert(0, ".")
exec(open("scripts/shim_latency.py").read())
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    one()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:7000])

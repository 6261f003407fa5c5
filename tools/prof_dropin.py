"""cProfile of the drop-in evaluation loop (bench.py --api dropin): where the host time of a batch goes."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ['bench.py', '--api', 'dropin', '--workload', sys.argv[1] if len(sys.argv) > 1 else 'reddit', '--steps', '300',
            '--warmup', '20', '--cpu-batches', '0']
import bench
pr = cProfile.Profile()
pr.enable()
bench.main()
pr.disable()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats('cumulative').print_stats(45)
print(out.getvalue()[:9000], file=sys.stderr)

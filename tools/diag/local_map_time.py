"""Times one local map (bench.local_map_bench) per regime; RSS_NO_POINT_SORT=1 gives the generic path for comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
with rss.Context(rss.DEFAULT_CONFIG, None, 0) as ctx:
    for wxyz, wrgb, name in ((0.5, 4.0, "node scales"), (20.0, 40.0, "fine scales")):
        r = bench.local_map_bench(ctx, synth, N, wxyz, wrgb, 6531.6, name)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})

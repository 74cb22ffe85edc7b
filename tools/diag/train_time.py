"""Times rss_forest_train on the bench's training set (no CPU arm)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
import rovinasemanticsegmentation_b200 as rss
from rovinasemanticsegmentation_b200 import synth
with rss.Context(rss.DEFAULT_CONFIG, bench.FOREST if hasattr(bench, "FOREST") else None, 0) as ctx:
    print(bench.forest_train_bench(ctx, synth, False))

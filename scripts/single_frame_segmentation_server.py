#!/usr/bin/env python
"""Drop-in for the reference's scripts/single_frame_segmentation_server.py: same node name, same service, answered by the
CUDA frame path (rovinasemanticsegmentation_b200/service.py)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from rovinasemanticsegmentation_b200.service import main

if __name__ == "__main__":
    main()

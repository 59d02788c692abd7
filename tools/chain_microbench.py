"""GPU micro-benchmark: a chain of 48 dependent launches of ONE TMA convolution inside a CUDA graph (ping-pong between two planes
buffers, so launch i + 1 reads what launch i wrote) -> microseconds per link, warm.  The product's batch-1 step is such a chain;
tools/launch_floor.cu gives the floor of the launch shape itself (2.2 us per link)."""
import ctypes, json, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from clslam_b200 import _lib as L
from tests import cabi
from tools.chain_microbench_lib import chain

cases = [(1, 48, 160, 64, L.PAD_ZERO, 3, True), (1, 24, 80, 128, L.PAD_ZERO, 3, True), (1, 24, 80, 128, L.PAD_ZERO, 3, False),
         (1, 12, 40, 256, L.PAD_ZERO, 3, True), (1, 12, 40, 256, L.PAD_ZERO, 3, False), (1, 6, 20, 512, L.PAD_ZERO, 3, True),
         (2, 48, 160, 64, L.PAD_ZERO, 3, True), (2, 24, 80, 128, L.PAD_ZERO, 3, True), (2, 6, 20, 512, L.PAD_ZERO, 3, True),
         (1, 24, 80, 128, L.PAD_ZERO, 1, False), (1, 96, 320, 32, L.PAD_REFLECT, 3, False), (1, 192, 640, 16, L.PAD_REFLECT, 3, False),
         (8, 48, 160, 64, L.PAD_ZERO, 3, True), (8, 24, 80, 128, L.PAD_ZERO, 3, True), (8, 12, 40, 256, L.PAD_ZERO, 3, True)]
for c in cases:
    try:
        print(json.dumps(chain(*c)), flush=True)
    except Exception as e:
        print(json.dumps({'case': str(c), 'error': str(e)[:200]}), flush=True)

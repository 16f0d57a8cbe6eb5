import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from tools.attn_tc_check import run
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
run(128, 32, 8, [T], [T], check=False, iters=2)

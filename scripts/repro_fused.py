import sys, os
sys.path.insert(0, os.getcwd())
from tests import util
out = util.run_cuda(160, 96, 2, mode="fused", keep=("weights","result"))
print("ok", out[-1]["weights"].shape)

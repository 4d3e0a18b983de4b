import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from test_gpu_parity import load, make_trainer
from multi_modal_normative_modeling_b200 import _lib
g = load('/root/repo/tests/golden', 'mm_M1_D116_full')
tr, _ = make_trainer(g)
print('engine', tr.engine(0), tr.engine(_lib.TRAIN_FP32), tr.engine(_lib.TRAIN_TC_SIMPLE))

import os, sys
sys.path.insert(0, os.getcwd())
import torch
from openpystruct_b200 import frames
p = frames.FrameOptParams(num_epochs=60, early_stop=False)
dev = torch.device("cuda", 0)
b = [(10, 10)] * 296
nb = torch.tensor([f[0] for f in b], dtype=torch.int32, device=dev); ns = torch.tensor([f[1] for f in b], dtype=torch.int32, device=dev)
frames.optimise_frames_device(p, nb, ns); torch.cuda.synchronize()

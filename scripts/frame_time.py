import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random, torch
from openpystruct_b200 import frames
p = frames.FrameOptParams(num_epochs=300, early_stop=False)
rng = random.Random(0)
batch = [frames.draw_frame(p, rng) for _ in range(592)]
dev = torch.device("cuda", 0)
def run(b, reps):
    nb = torch.tensor([f[0] for f in b], dtype=torch.int32, device=dev); ns = torch.tensor([f[1] for f in b], dtype=torch.int32, device=dev)
    frames.optimise_frames_device(p, nb, ns); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = frames.optimise_frames_device(p, nb, ns)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = run(batch, 3); print("592 random frames: frames/s", 592 / ms * 1e3, "ms", ms)
ms = run([(10, 10)] * 296, 1); print("296 frames of 10x10: ms", ms, "us per epoch", ms * 1e3 / 300)
ms = run([(5, 5)] * 296, 1); print("296 frames of 5x5: ms", ms, "us per epoch", ms * 1e3 / 300)

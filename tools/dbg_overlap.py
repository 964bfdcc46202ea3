"""Debug: bench-shaped overlapped steps (CNN of batch k+1 next to the solve of batch k), eager launches."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from bench import build_network, synthetic_batch
from crossloc_b200.pipeline import Localizer
dev = torch.device('cuda', 0)
net = build_network(dev)
loc = Localizer(net, hyps=256, device=dev)
images, offsets, _, focal = synthetic_batch(0, 32)
images, offsets, focal = images.to(dev), offsets.to(dev), focal.to(dev)
for s in range(6):
    if s == 3 and os.environ.get('DBG_PROFILE', '1') == '1':
        net._runtime.set_profiling(True)
    loc.localize_device(images, focal, offsets, image_base=32 * s, overlap=True)
    if os.environ.get('DBG_FLUSH', '0') == '1' or s in (2, 5):
        loc.flush()
        torch.cuda.synchronize()
    print('step', s, 'ok', flush=True)

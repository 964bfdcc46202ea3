"""Where a batch-1 frame spends its time (the reference's call pattern, BASELINE config 1)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import HEIGHT, WIDTH, build_network  # noqa: E402

dev = torch.device('cuda', 0)
net = build_network(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x_host = torch.rand(B, 3, HEIGHT, WIDTH)
x_pin = x_host.pin_memory()
x_dev = x_host.to(dev)


def wall(fn, n=100, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


with torch.no_grad():
    print('forward on a device tensor, back to back (graph): %.3f ms' % wall(lambda: net(x_dev)))
    print('forward + sync each:                              %.3f ms' % wall(lambda: (net(x_dev), torch.cuda.synchronize())))
    print('pageable H2D of the frame + sync:                 %.3f ms' % wall(lambda: (x_host.cuda(), torch.cuda.synchronize())))
    print('pinned H2D of the frame + sync:                   %.3f ms' % wall(lambda: (x_pin.cuda(), torch.cuda.synchronize())))
    out = net(x_dev)
    print('.cpu() of the coordinate map:                     %.3f ms' % wall(lambda: out[:, :3].cpu()))
    print('network(image.cuda()) ... .cpu():                 %.3f ms' % wall(lambda: net(x_host.cuda())[:, :3].cpu()))
    rt = net._runtime
    rt.set_profiling(True)
    for _ in range(3):
        net(x_dev)
    torch.cuda.synchronize()
    rt.read_profile(B, HEIGHT, WIDTH)
    for _ in range(10):
        net(x_dev)
    torch.cuda.synchronize()
    prof = rt.read_profile(B, HEIGHT, WIDTH, keep_enabled=False)
    agg = {}
    for kind, label, flops, ms, forwards in prof:
        a = agg.setdefault((kind, str(label)), [0, 0.0])
        a[0] += 1
        a[1] += ms / forwards
    tot = sum(v[1] for v in agg.values())
    print('per-op device time (eager, events between ops), %.3f ms per forward:' % tot)
    for (kind, label), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print('   %-10s %-28s x%-3d %.3f ms' % (kind, label, n, ms))

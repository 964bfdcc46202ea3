"""Experiment: one forward of 32 frames vs two concurrent forwards of 16 frames on two streams (two handles), so that the
memory-bound passes of one lane overlap the tensor-bound convolutions of the other."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import networks.networks as nets  # noqa: E402

torch.manual_seed(2021)
net = nets.TransPoseNet(torch.zeros(3), False, False, 2, 2, 3, 1).eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x = torch.rand(B, 3, 480, 720, device='cuda')


def timed(fn, n=10, warm=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    ref = net(x).clone()
    print('one lane, %d frames: %.2f ms' % (B, timed(lambda: net(x))), flush=True)
    nets_l = [net] + [copy.deepcopy(net) for _ in range(lanes - 1)]
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    xs = list(x.chunk(lanes))
    xs = [c.contiguous() for c in xs]
    outs = [None] * lanes
    main = torch.cuda.current_stream()

    def step():
        for i in range(lanes):
            streams[i].wait_stream(main)
            with torch.cuda.stream(streams[i]):
                outs[i] = nets_l[i](xs[i])
        for i in range(lanes):
            main.wait_stream(streams[i])

    ms = timed(step)
    print('%d lanes of %d frames: %.2f ms' % (lanes, B // lanes, ms), flush=True)
    got = torch.cat(outs)
    print('max abs diff vs one lane: %.3g (max |ref| %.3g)' % ((got - ref).abs().max().item(), ref.abs().max().item()))

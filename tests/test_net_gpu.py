"""The C++ network runtime (csrc/net.cu, cl_net_create / cl_net_forward) against the Python plan over the same kernels,
against plain fp32 torch, and through its raw C ABI with host buffers."""
import ctypes

import numpy as np
import pytest
import torch

import networks.networks as nets
from crossloc_b200 import _lib, net as native_net
from crossloc_b200.cnn import CoordNetEngine
from tests.test_cnn_cpu import build_case

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize('name', ['transpose_default', 'transpose_ragged', 'transpose_tiny_gray', 'network_vanilla',
                                  'network_tiny', 'transpose_fullsize_ragged', 'transpose_fullsize_even'])
def test_runtime_equals_python_plan(name, monkeypatch):
    """Same kernels, same order, same operand planes: the two hosts agree to GroupNorm-statistics round-off (fp64 atomics
    in a different order), eager launch and graph replay alike.  (fp16 + fp8 mode: the Python plan has no e2m1 path.)"""
    from crossloc_b200 import cnn
    monkeypatch.setattr(cnn, 'PRECISION', 'fp16+fp8')
    net, x = build_case(name, DEV)
    spec = net._spec(x) if isinstance(net, nets.TransPoseNet) else net._spec()
    assert native_net.supported(spec)
    with torch.no_grad():
        py = CoordNetEngine().forward(spec, x)
        eager = net(x)          # first call of a plan: direct launches
        graph = net(x)          # second call: captured graph
        graph2 = net(x)
    assert net._runtime is not None and net._runtime.launches_per_forward > 20
    assert eager.shape == py.shape
    assert rel_l2(eager, py) < 2e-6
    assert rel_l2(graph, py) < 2e-6
    assert rel_l2(graph2, graph) < 2e-6
    assert torch.isfinite(graph).all()


@pytest.mark.parametrize('name', ['transpose_default', 'transpose_ragged', 'network_vanilla', 'transpose_fullsize_even'])
@pytest.mark.parametrize('precision', ['fp16+fp4', 'fp16+fp8', 'fp16x3'])
def test_runtime_precisions_against_fp32_torch(name, precision, monkeypatch):
    """Every arithmetic scheme of the runtime against plain fp32 torch; fp16 + fp4 (block-scaled e2m1 corrections, the
    default) must keep a 3x margin on the 1e-3 bar."""
    from crossloc_b200 import cnn
    monkeypatch.setattr(cnn, 'PRECISION', precision)
    net, x = build_case(name, DEV)
    with torch.no_grad():
        out = net(x)
        out2 = net(x)
        ref = net.forward_reference(x)
    assert net._runtime is not None and net._runtime.precision == precision
    err = rel_l2(out[:, :3], ref[:, :3])
    assert err < {'fp16+fp4': 3.3e-4, 'fp16+fp8': 1e-4, 'fp16x3': 2e-5}[precision], err
    assert rel_l2(out2, out) < 2e-6


def test_runtime_follows_parameter_updates_and_new_sizes():
    torch.manual_seed(3)
    net = nets.TransPoseNet(torch.tensor([1.0, 2.0, 3.0]), True, False, 1, 1, 3, 1).eval().to(DEV)
    x = torch.rand(2, 3, 48, 64, device=DEV)
    with torch.no_grad():
        a = net(x)
        a2 = net(x)
        handle = net._runtime._handle.value
        # in-place update (optimizer step / load_state_dict): same storage, new values -> cl_net_update
        for p in net.parameters():
            p.mul_(1.01)
        net.decoder.mean.add_(5.0)
        b = net(x)
        ref_b = net.forward_reference(x)
        assert net._runtime._handle.value == handle
        assert rel_l2(b[:, :3], ref_b[:, :3]) < 1e-3 and rel_l2(a2, a) < 2e-6 and rel_l2(b, a) > 1e-3
        # another frame size: a second cached plan; the first one is still valid afterwards
        y = torch.rand(1, 3, 40, 72, device=DEV)
        c = net(y)
        assert rel_l2(c[:, :3], net.forward_reference(y)[:, :3]) < 1e-3
        assert rel_l2(net(x), b) < 2e-6
        # storage moves (module.to / new tensors): a fresh handle
        net.decoder.fc3.weight.data = net.decoder.fc3.weight.data.clone()
        d = net(x)
        assert rel_l2(d, b) < 2e-6


def test_runtime_frames_entry_matches_float_entry():
    """uint8 HWC frames (device and pinned host) through cl_net_forward_frames == ToTensor [+ Normalize] + cl_net_forward."""
    torch.manual_seed(4)
    net = nets.TransPoseNet(torch.zeros(3), True, False, 0, 1, 3, 1).eval().to(DEV)
    frames = torch.randint(0, 256, (2, 48, 80, 3), dtype=torch.uint8)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    x = frames.permute(0, 3, 1, 2).float().div(255.0)
    xn = (x - torch.tensor(mean)[None, :, None, None]) / torch.tensor(std)[None, :, None, None]
    with torch.no_grad():
        for _ in range(3):   # eager, then graph replays
            raw_dev = net.forward_frames(frames.to(DEV))
            raw_host = net.forward_frames(frames.pin_memory())
        want = net(x.to(DEV))
        assert rel_l2(raw_dev, want) < 2e-6 and rel_l2(raw_host, want) < 2e-6
        net2 = nets.TransPoseNet(torch.zeros(3), True, False, 0, 1, 3, 1).eval().to(DEV)
        net2.load_state_dict(net.state_dict())
        got = net2.forward_frames(frames.to(DEV), mean, std)
        assert rel_l2(got, net2(xn.to(DEV))) < 2e-6


@pytest.mark.parametrize('co_task,co_pos', [(3, 0), (3, 1), (6, 0), (6, 1)])
def test_full_size_head_any_channel_count(co_task, co_pos):
    """The DUC GroupNorm has 2 * Co channels per group (6, 8, 12, 14): sizes the convolution epilogue does not reduce are
    summed by a separate statistics kernel (semantics / uncertainty-free variants used to raise)."""
    torch.manual_seed(20 + co_task + co_pos)
    net = nets.TransPoseNet(torch.zeros(co_task), True, False, 0, 0, co_task, co_pos, full_size_output=True).eval().to(DEV)
    x = torch.rand(1, 3, 40, 56, device=DEV)
    with torch.no_grad():
        out = net(x)
        out2 = net(x)
        ref = net.forward_reference(x)
    assert out.shape == ref.shape == (1, co_task + co_pos, 40, 56)
    assert rel_l2(out[:, :co_task], ref[:, :co_task]) < 1e-3
    assert rel_l2(out2, out) < 2e-6


def test_encoder_and_decoder_alone_run_native():
    """TransPoseNetEncoder.forward / TransPoseNetDecoder.forward (the reference composes them per MLR branch,
    networks.py:484-500) run the native plan, not stock cuDNN."""
    torch.manual_seed(6)
    net = nets.TransPoseNet(torch.tensor([1.0, -1.0, 2.0]), True, False, 1, 1, 3, 1).eval().to(DEV)
    x = torch.rand(2, 3, 48, 64, device=DEV)
    lib = _lib.load()
    with torch.no_grad():
        feat = net.encoder(x)
        feat_ref = net.encoder.forward_reference(x)
        assert net.encoder._engine is not None and net.encoder._engine.launches > 10
        assert rel_l2(feat, feat_ref) < 3.3e-4
        out = net.decoder(feat_ref)
        out_ref = net.decoder.forward_reference(feat_ref)
        assert net.decoder._engine is not None and net.decoder._engine.launches > 5
        assert rel_l2(out[:, :3], out_ref[:, :3]) < 3.3e-4
    assert lib is not None


def test_c_abi_with_host_buffers():
    """cl_net_forward straight through ctypes with pageable host memory on both sides, as a C++ host would call it."""
    torch.manual_seed(8)
    net = nets.TransPoseNet(torch.zeros(3), True, False, 0, 0, 3, 1).eval().to(DEV)
    x = torch.rand(1, 3, 32, 48)
    with torch.no_grad():
        want = net(x.to(DEV)).cpu()
    rt = net._runtime
    lib = _lib.load()
    c, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.cl_net_output_shape(rt._handle, 1, 32, 48, ctypes.byref(c), ctypes.byref(h), ctypes.byref(w)))
    assert (c.value, h.value, w.value) == (4, 4, 6)
    out = np.zeros((1, 4, 4, 6), dtype=np.float32)
    xin = np.ascontiguousarray(x.numpy())
    _lib.check(lib.cl_net_forward(rt._handle, xin.ctypes.data, 1, 32, 48, out.ctypes.data, None))
    assert np.abs(out - want.numpy()).max() <= 2e-6 * np.abs(want.numpy()).max()
    # errors are reported, not swallowed
    assert lib.cl_net_forward(rt._handle, xin.ctypes.data, 0, 32, 48, out.ctypes.data, None) != 0
    assert b'invalid sizes' in lib.cl_last_error()


def test_profile_mode_lists_every_launch():
    torch.manual_seed(9)
    net = nets.TransPoseNet(torch.zeros(3), True, False, 0, 0, 3, 1).eval().to(DEV)
    x = torch.rand(2, 3, 32, 48, device=DEV)
    with torch.no_grad():
        base = net(x)
        net._runtime.set_profiling(True)
        for _ in range(3):
            prof_out = net(x)
        rows = net._runtime.read_profile(2, 32, 48, keep_enabled=False)
        again = net(x)
    assert rel_l2(prof_out, base) < 2e-6 and rel_l2(again, base) < 2e-6
    kinds = [r[0] for r in rows]
    assert kinds.count('stem') == 2 and kinds.count('head') == 1 and kinds.count("conv") == 14
    assert all(r[4] == 3 for r in rows) and sum(r[3] for r in rows) > 0
    convs = [r for r in rows if r[0] == 'conv']
    assert convs[0][1] == (32, 64, 3, 2) and convs[0][2] == 2.0 * 2 * 16 * 24 * 64 * 32 * 9


def _fresh_runtime_output(net, x, env, precision='fp16+fp8'):
    """Forward through a NEW cl_net handle created under the given environment switches (read at cl_net_create).  The
    fused epilogue only exists for the fp16 + fp8 scheme, hence the pinned precision."""
    import os
    from crossloc_b200 import cnn
    saved = {k: os.environ.get(k) for k in env}
    saved_precision = cnn.PRECISION
    os.environ.update(env)
    cnn.PRECISION = precision
    try:
        net._runtime = None
        with torch.no_grad():
            outs = [net(x) for _ in range(3)]   # eager, graph, graph
    finally:
        cnn.PRECISION = saved_precision
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        net._runtime = None
    for o in outs[1:]:
        assert rel_l2(o, outs[0]) < 2e-6
    return outs[0]


@pytest.mark.parametrize('shape', [(2, 3, 256, 384), (1, 3, 480, 720), (3, 3, 200, 312)])
def test_fused_groupnorm_epilogue_equals_two_pass_path(shape):
    """Sizes where the CTA-pair kernel runs: GroupNorm / ReLU / residual merge in the convolution epilogue (accumulators
    wait in tensor memory for the image statistics, dynamic tile scheduler) vs the raw tensor + gn_apply passes, with
    static and with dynamic tile scheduling, and vs plain fp32 torch."""
    torch.manual_seed(31)
    net = nets.TransPoseNet(torch.tensor([1.0, -2.0, 0.5]), False, False, 1, 1, 3, 1).eval().to(DEV)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(7)).to(DEV)
    fused = _fresh_runtime_output(net, x, {'CROSSLOC_B200_FUSE_GN': '1', 'CROSSLOC_B200_DYNAMIC_TILES': '1'})
    two_pass_dyn = _fresh_runtime_output(net, x, {'CROSSLOC_B200_FUSE_GN': '0', 'CROSSLOC_B200_DYNAMIC_TILES': '1'})
    two_pass_static = _fresh_runtime_output(net, x, {'CROSSLOC_B200_FUSE_GN': '0', 'CROSSLOC_B200_DYNAMIC_TILES': '0'})
    with torch.no_grad():
        ref = net.forward_reference(x)
    assert rel_l2(two_pass_dyn, two_pass_static) < 2e-6
    assert rel_l2(fused, two_pass_static) < 5e-6
    assert rel_l2(fused[:, :3], ref[:, :3]) < 3.3e-4


def test_fused_epilogue_vanilla_network_and_many_small_images():
    """No GroupNorm (vanilla Network: nothing to wait for) and a batch whose images are smaller than a tile (a 128-row
    tile spans several images: per-image publication counts)."""
    torch.manual_seed(32)
    net = nets.Network(torch.tensor([1.0, 2.0, 3.0]), False).eval().to(DEV)
    x = torch.rand(2, 1, 256, 320, device=DEV)
    fused = _fresh_runtime_output(net, x, {'CROSSLOC_B200_FUSE_GN': '1'})
    plain = _fresh_runtime_output(net, x, {'CROSSLOC_B200_FUSE_GN': '0'})
    with torch.no_grad():
        ref = net.forward_reference(x)
    assert rel_l2(fused, plain) < 5e-6 and rel_l2(fused, ref) < 3.3e-4
    torch.manual_seed(33)
    net2 = nets.TransPoseNet(torch.zeros(3), False, False, 0, 1, 3, 1).eval().to(DEV)
    y = torch.rand(40, 3, 48, 64, device=DEV)      # 6 x 8 cells: padded plane of 80 rows, 40 images -> 25 tiles
    fused = _fresh_runtime_output(net2, y, {'CROSSLOC_B200_FUSE_GN': '1'})
    plain = _fresh_runtime_output(net2, y, {'CROSSLOC_B200_FUSE_GN': '0'})
    with torch.no_grad():
        ref = net2.forward_reference(y)
    assert rel_l2(fused, plain) < 5e-6 and rel_l2(fused[:, :3], ref[:, :3]) < 3.3e-4


def test_forwards_of_one_handle_on_two_streams_are_ordered():
    """The plans of a handle share their input / activation / output buffers: forwards issued on different streams are ordered
    by the handle's own event (forwards of different handles may overlap), so alternating streams gives the single-stream maps."""
    torch.manual_seed(8)
    net = nets.TransPoseNet(torch.zeros(3), False, False, 1, 1, 3, 1).eval().to(DEV)
    xs = [torch.rand(2, 3, 96, 144, device=DEV) for _ in range(4)]
    with torch.no_grad():
        want = [net(x).clone() for x in xs]
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        got = []
        for rounds in range(3):
            got = []
            for i, x in enumerate(xs):
                s = streams[i % 2]
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    got.append(net(x).clone())
        torch.cuda.synchronize()
    for a, b in zip(got, want):
        assert rel_l2(a, b) < 2e-6

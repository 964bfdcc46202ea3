"""CPU checks of the host-side logic around the kernels: tap tables of the padded-flat layout, the per-layer arithmetic
policy, sharded geometry helpers and the profile summariser.  Nothing here touches a GPU."""
import os
import subprocess
import sys

import torch

from crossloc_b200 import cnn, layout
from crossloc_b200.cnn import _Geometry, _nterms_for, _taps
from tests.util import ROOT


class _Pack:
    def __init__(self, ksize, stride):
        self.ksize, self.stride = ksize, stride


def _conv_by_taps(x, weight, stride):
    """Convolution evaluated exactly as the kernels do: sum over taps of row-shifted GEMMs on the padded-flat matrix."""
    b, cin, h, w = x.shape
    cout = weight.size(0)
    ho, wo = ((h + 1) // 2, (w + 1) // 2) if stride == 2 else (h, w)
    geo = _Geometry(b, ho, wo)
    phases = 4 if stride == 2 else 1
    act = layout.to_pf(x, phases=phases, terms=1).to(torch.float64)          # [phases * Mp][cin]
    taps = _taps(_Pack(weight.size(2), stride), geo)
    wt = weight.permute(2, 3, 0, 1).reshape(-1, cout, cin).to(torch.float64)  # [tap][cout][cin]
    rows = torch.arange(geo.Mp)
    out = torch.zeros(geo.Mp, cout, dtype=torch.float64)
    for t, shift in enumerate(taps):
        src = rows + shift
        ok = (src >= 0) & (src < act.size(0))
        a = torch.zeros(geo.Mp, cin, dtype=torch.float64)
        a[ok] = act[src[ok]]
        out += a @ wt[t].T
    return layout.raw_to_nchw(out.to(torch.float32), b, ho, wo)


def test_tap_tables_reproduce_convolutions():
    """A filter tap is a constant row shift in the padded-flat layout (stride 1) / a phase plane plus a shift (stride 2)."""
    torch.manual_seed(0)
    for (k, stride, h, w) in ((3, 1, 6, 7), (1, 1, 5, 4), (3, 2, 8, 10), (3, 2, 7, 9)):
        x = torch.randn(2, 8, h, w).half().float()      # fp16-exact inputs: the layout conversion is then lossless
        weight = torch.randn(16, 8, k, k)
        ref = torch.nn.functional.conv2d(x, weight, None, stride, k // 2)
        got = _conv_by_taps(x, weight, stride)
        assert got.shape == ref.shape
        assert float((got - ref).abs().max()) < 1e-4


def test_tap_phase_split_used_by_the_weight_gradient():
    """Splitting a stride-2 tap row into (phase plane, in-plane shift) as train_plan / train do."""
    geo = _Geometry(3, 9, 11)
    for t in _taps(_Pack(3, 2), geo):
        ph = (t + geo.Mp // 2) // geo.Mp
        shift = t - ph * geo.Mp
        assert 0 <= ph < 4 and -(geo.Wp + 1) <= shift <= 0


def test_arithmetic_policy_per_layer():
    assert _nterms_for('fp16+fp8', 512, 3, 1) == 2 and _nterms_for('fp16+fp8', 256, 3, 1) == 2
    assert _nterms_for('fp16+fp8', 128, 3, 1) == 3           # too narrow for the e4m3 planes
    assert _nterms_for('fp16+fp8', 128, 3, 2) == 3 and _nterms_for('fp16+fp8', 32, 3, 2) == 3
    assert _nterms_for('fp16+fp8', 512, 1, 1) == (2 if cnn._FP8_1X1 else 3)
    assert _nterms_for('fp16+fp8', 256, 1, 1) == 3
    assert _nterms_for('fp16x3', 512, 3, 1) == 3 and _nterms_for('fp16x1', 512, 3, 1) == 1


def test_pf_layout_round_trip():
    x = torch.randn(2, 8, 5, 7)
    assert torch.allclose(layout.from_pf(layout.to_pf(x), 2, 5, 7), x, atol=1e-6)
    assert torch.allclose(layout.from_pf_phases(layout.to_pf(x, phases=4), 2, 5, 7), x, atol=1e-6)
    pf = layout.to_pf(x)
    rows = 2 * 7 * 9
    assert pf.shape == (2 * rows, 8)
    border = pf[:rows].reshape(2, 7, 9, 8)
    assert float(border[:, 0].abs().max()) == 0 and float(border[:, :, 0].abs().max()) == 0     # zero borders


def test_profile_summariser_steady_state(tmp_path):
    csv_path = tmp_path / 'launches.csv'
    rows = ['"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC",'
            '"Section Name","Metric Name","Metric Unit","Metric Value"']
    names = ['setup', 'void cl::head_kernel<4>(x)', 'void cl::conv(x)', 'void cl::head_kernel<4>(x)', 'void cl::conv(x)',
             'void cl::conv(x)', 'void cl::head_kernel<4>(x)']
    for i, n in enumerate(names):
        rows.append('"%d","1","p","h","%s","1","7","(1, 1, 1)","(1, 1, 1)","0","10.0","s","gpu__time_duration.sum","ns","1000"' % (i, n))
    csv_path.write_text('\n'.join(rows) + '\n')
    out = tmp_path / 'summary.md'
    subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'summarize_profiles.py'), 'launches', str(csv_path), '3', str(out),
                    'title', 'head_kernel'], check=True)
    text = out.read_text()
    assert '| **total** | 5 | 2.5 |' in text      # the launches after the first marker up to the last one, over two steps


def test_precision_decision_by_cpu_emulation():
    """tools/emulate_precision.py: rounding the convolution operands once (fp16 / TF32) sits at the 1e-3 parity bar, the
    shipped fp16 + e4m3 scheme two orders below it, and block-scaled FP4 corrections in between."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import emulate_precision
    errs = emulate_precision.run(32, 48, 1)
    e = {k: v[0] for k, v in errs.items()}
    assert e['fp16x3'] < 2e-6
    assert e['fp16x3'] < e['fp16+fp8'] < 1e-4
    assert e['fp16+fp8'] < e['fp16+fp4'] < 5e-4
    assert e['fp16+fp4'] < e['fp16+fp8a'] and e['fp16+fp4'] < e['fp16+fp8w']
    assert 3e-4 < e['fp16x1'] < 5e-3

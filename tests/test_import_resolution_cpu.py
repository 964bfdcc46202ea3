"""The documented switch-over (this repository on PYTHONPATH, scripts run from the CrossLoc checkout)
must bind `networks.networks`, `dsacstar` and `loss.coord` to the twins here while `loss.depth` / `loss.normal` /
`loss.semantics` (imported unconditionally by /root/reference/train_single_task.py:12-15) still resolve to the checkout."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROBE = textwrap.dedent('''
    import importlib.util, json, sys
    import torch
    from loss.coord import get_cam_mat, scene_coords_regression_loss
    from loss.depth import depth_regression_loss
    from loss.normal import normal_regression_loss
    import networks.networks as nets
    import dsacstar
    print(json.dumps({
        'coord': sys.modules['loss.coord'].__file__, 'depth': sys.modules['loss.depth'].__file__,
        'normal': sys.modules['loss.normal'].__file__, 'networks': nets.__file__, 'dsacstar': dsacstar.__file__,
        'native': hasattr(nets.TransPoseNet, 'forward_reference') and callable(dsacstar.forward_rgb),
        'path0': sys.path[0]}))
''')


def _fake_checkout(tmp_path):
    """A directory shaped like the CrossLoc checkout: namespace directories without __init__.py."""
    for sub in ('loss', 'networks', 'dsacstar'):
        (tmp_path / sub).mkdir()
    (tmp_path / 'loss' / 'coord.py').write_text('def get_cam_mat(*a):\n    return "checkout"\n'
                                                'def scene_coords_regression_loss(*a):\n    return "checkout"\n')
    (tmp_path / 'loss' / 'depth.py').write_text('def depth_regression_loss(*a):\n    return "checkout"\n')
    (tmp_path / 'loss' / 'normal.py').write_text('def normal_regression_loss(*a):\n    return "checkout"\n')
    (tmp_path / 'networks' / 'networks.py').write_text('class TransPoseNet:\n    pass\n')
    (tmp_path / 'dsacstar' / 'dsacstar.cpp').write_text('// sources only, as in the checkout\n')
    (tmp_path / 'train_probe.py').write_text(PROBE)
    return tmp_path


def test_twins_win_and_other_losses_still_resolve(tmp_path):
    import json
    checkout = _fake_checkout(tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT)
    # run as a script from the checkout: its directory becomes sys.path[0], ahead of PYTHONPATH
    out = subprocess.run([sys.executable, str(checkout / 'train_probe.py')], cwd=str(checkout), env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads(out.stdout.strip().splitlines()[-1])
    assert os.path.samefile(info['path0'], str(checkout))
    assert info['coord'] == os.path.join(ROOT, 'loss', 'coord.py')
    assert info['networks'] == os.path.join(ROOT, 'networks', 'networks.py')
    assert info['dsacstar'] == os.path.join(ROOT, 'dsacstar', '__init__.py')
    assert info['depth'] == str(checkout / 'loss' / 'depth.py')
    assert info['normal'] == str(checkout / 'loss' / 'normal.py')
    assert info['native'] is True


def test_resolution_against_the_real_checkout_when_present():
    ref = '/root/reference'
    if not os.path.isdir(os.path.join(ref, 'loss')):
        import pytest
        pytest.skip('reference checkout not present (GPU box)')
    code = ('import sys, importlib.util as u; sys.path[:0] = [%r, %r]; import torch; '
            'print(u.find_spec("loss.depth").origin); print(u.find_spec("loss.coord").origin); '
            'print(u.find_spec("networks.networks").origin)' % (ref, ROOT))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    depth, coord, nets = out.stdout.strip().splitlines()[-3:]
    assert depth == os.path.join(ref, 'loss', 'depth.py')
    assert coord == os.path.join(ROOT, 'loss', 'coord.py')
    assert nets == os.path.join(ROOT, 'networks', 'networks.py')

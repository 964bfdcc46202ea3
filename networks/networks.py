"""Drop-in twin of the reference's `networks.networks` for the localization hot path.

Same classes, constructor signatures, attributes and state-dict keys as
/root/reference/networks/networks.py (`Network` :43-130, `TransPoseNetEncoder` :175-256,
`TransPoseNetDecoder` :276-360, `TransPoseNet` :363-502), so `load_state_dict(strict=True)` of reference
checkpoints and the unchanged callers (utils/evaluation.py:106-116, test_single_task.py:347-366) keep working.
What differs is what `forward` launches: on a CUDA tensor without autograd it runs the hand-written sm_100a
kernels of crossloc_b200 (TMA + tcgen05 implicit-GEMM convolutions, fused GroupNorm statistics) instead of
cuDNN/ATen; with autograd enabled (`train_single_task.py:262-299`) it runs the fused training plan
(`crossloc_b200.train_plan`: one autograd node whose backward is GroupNorm / ReLU backward, data and weight
gradient kernels).  `forward_reference` is the same network spelled with stock torch ops in fp32: the definition
the parity tests compare against.  There is no CPU execution of the native path.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from crossloc_b200 import net as native_net
from crossloc_b200 import train as native_train
from crossloc_b200 import train_plan
from crossloc_b200.cnn import CoordNetEngine

try:   # the reference logs through utils.io.safe_printout (networks.py:40); optional here
    from utils.io import safe_printout
except Exception:   # pragma: no cover - reference utils not on the path
    def safe_printout(words):
        pass

_FUSED_TRAIN = os.environ.get('CROSSLOC_B200_TRAIN', 'fused') != 'layerwise'
_POS_CLAMP = (-16.10, 13.82)   # exp() range limiter of the uncertainty channel, networks.py:355-356


def _width(tiny, full=512):
    return 128 if tiny else full


def _res_block(tiny, num_gn_channel, in_ch=None):
    """conv3x3-GN-ReLU, conv1x1-GN-ReLU, conv3x3-GN-ReLU (networks.py:133-146; :149-163 for in_ch != width)."""
    ch = _width(tiny)
    in_ch = ch if in_ch is None else in_ch
    groups = min(num_gn_channel, ch)
    return nn.Sequential(
        nn.Conv2d(in_ch, ch, 3, 1, 1), nn.GroupNorm(groups, ch), nn.ReLU(),
        nn.Conv2d(ch, ch, 1, 1, 0), nn.GroupNorm(groups, ch), nn.ReLU(),
        nn.Conv2d(ch, ch, 3, 1, 1), nn.GroupNorm(groups, ch), nn.ReLU())


def _torch_conv(conv, x):
    return conv(x)


def _run_block(block, x, conv):
    """nn.Sequential of conv / GroupNorm / ReLU with a pluggable convolution operator."""
    for m in block:
        x = conv(m, x) if isinstance(m, nn.Conv2d) else m(x)
    return x


def _run_native(module, spec, inputs):
    """Inference on the native kernels: the C++ runtime (cl_net_forward: cached plan, CUDA graph) for every plan it covers,
    else -- or with CROSSLOC_B200_ENGINE=python -- the Python plan of crossloc_b200.cnn over the same kernels."""
    if native_net.ENGINE == 'native' and native_net.supported(spec):
        if module._runtime is None:
            module._runtime = native_net.NetRuntime()
        return module._runtime.forward(spec, inputs)
    if module._engine is None:
        module._engine = CoordNetEngine()
    return module._engine.forward(spec, inputs)


def _native_ok(module, x):
    if not x.is_cuda:
        raise RuntimeError('crossloc_b200: forward() needs a CUDA tensor -- the native path has no CPU fallback '
                           '(use forward_reference() for the plain-torch definition)')
    return not (torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()))


class Network(nn.Module):
    """Vanilla DSAC* FCN (networks.py:43-130): grayscale in, 3-channel scene coordinates out, no normalisation."""

    OUTPUT_SUBSAMPLE = 8

    def __init__(self, mean, tiny):
        super(Network, self).__init__()
        c4, c5 = _width(tiny, 256), _width(tiny, 512)
        self.conv1 = nn.Conv2d(1, 32, 3, 1, 1)
        self.conv2 = nn.Conv2d(32, 64, 3, 2, 1)
        self.conv3 = nn.Conv2d(64, 128, 3, 2, 1)
        self.conv4 = nn.Conv2d(128, c4, 3, 2, 1)
        self.res1_conv1 = nn.Conv2d(c4, c4, 3, 1, 1)
        self.res1_conv2 = nn.Conv2d(c4, c4, 1, 1, 0)
        self.res1_conv3 = nn.Conv2d(c4, c4, 3, 1, 1)
        self.res2_conv1 = nn.Conv2d(c4, c5, 3, 1, 1)
        self.res2_conv2 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.res2_conv3 = nn.Conv2d(c5, c5, 3, 1, 1)
        if not tiny:
            self.res2_skip = nn.Conv2d(256, 512, 1, 1, 0)
        self.res3_conv1 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.res3_conv2 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.res3_conv3 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.fc1 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.fc2 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.fc3 = nn.Conv2d(c5, 3, 1, 1, 0)
        self.register_buffer('mean', mean.clone())
        self.tiny = tiny
        self._engine = None
        self._runtime = None

    def _spec(self):
        names = ['conv1', 'conv2', 'conv3', 'conv4', 'res1_conv1', 'res1_conv2', 'res1_conv3', 'res2_conv1',
                 'res2_conv2', 'res2_conv3', 'res3_conv1', 'res3_conv2', 'res3_conv3', 'fc1', 'fc2']
        if not self.tiny:
            names.append('res2_skip')
        res2 = {'kind': 'residual', 'convs': ['res2_conv1', 'res2_conv2', 'res2_conv3']}
        if not self.tiny:
            res2 = {'kind': 'residual_skip', 'convs': res2['convs'], 'skip': 'res2_skip'}
        return {
            'group_norm': False,
            'layers': [(n, getattr(self, n), None) for n in names],
            'blocks': [{'kind': 'residual', 'convs': ['res1_conv1', 'res1_conv2', 'res1_conv3']}, res2,
                       {'kind': 'residual', 'convs': ['res3_conv1', 'res3_conv2', 'res3_conv3']},
                       {'kind': 'plain', 'convs': ['fc1', 'fc2']}],
            'head': {'conv': self.fc3, 'mean': self.mean, 'num_task': 3, 'clamp': _POS_CLAMP},
        }

    def forward_reference(self, inputs, conv=_torch_conv):
        x = F.relu(conv(self.conv1, inputs))
        x = F.relu(conv(self.conv2, x))
        x = F.relu(conv(self.conv3, x))
        res = F.relu(conv(self.conv4, x))
        x = F.relu(conv(self.res1_conv1, res))
        x = F.relu(conv(self.res1_conv2, x))
        x = F.relu(conv(self.res1_conv3, x))
        res = res + x
        x = F.relu(conv(self.res2_conv1, res))
        x = F.relu(conv(self.res2_conv2, x))
        x = F.relu(conv(self.res2_conv3, x))
        if not self.tiny:
            res = conv(self.res2_skip, res)
        res = res + x
        x = F.relu(conv(self.res3_conv1, res))
        x = F.relu(conv(self.res3_conv2, x))
        x = F.relu(conv(self.res3_conv3, x))
        res = res + x
        sc = F.relu(conv(self.fc1, res))
        sc = F.relu(conv(self.fc2, sc))
        sc = conv(self.fc3, sc)
        return sc + self.mean.to(sc.device)[None, :, None, None]

    def forward_train(self, inputs, fused=None):
        """Autograd path on the native kernels: the fused plan (crossloc_b200.train_plan: one autograd node, GroupNorm /
        ReLU / residual backward and both convolution gradients on padded-flat tensors) or, with fused=False, the
        per-layer path (conv forward / dgrad / wgrad as an autograd Function around stock pointwise ops)."""
        if fused is None:
            fused = _FUSED_TRAIN
        if fused and inputs.is_cuda:
            return train_plan.forward_train(self, inputs)
        return self.forward_reference(inputs, conv=native_train.conv2d)

    def forward(self, inputs):
        if not _native_ok(self, inputs):
            return self.forward_train(inputs)
        return _run_native(self, self._spec(), inputs)


class TransPoseNetEncoder(nn.Module):
    """Encoder (networks.py:175-256): strided conv ladder + two residual stages + optional extra blocks."""

    def __init__(self, tiny, grayscale, enc_add_res_block=0, num_gn_channel=32):
        super(TransPoseNetEncoder, self).__init__()
        self.tiny = tiny
        self.grayscale = grayscale
        self.enc_add_res_block = enc_add_res_block
        self.num_gn_channel = num_gn_channel
        c4, c5, g = _width(tiny, 256), _width(tiny, 512), num_gn_channel
        self.conv1 = nn.Conv2d(1 if grayscale else 3, g, 3, 1, 1)
        self.norm1 = nn.GroupNorm(g, g)
        self.conv2 = nn.Conv2d(g, 64, 3, 2, 1)
        self.norm2 = nn.GroupNorm(g, 64)
        self.conv3 = nn.Conv2d(64, 128, 3, 2, 1)
        self.norm3 = nn.GroupNorm(g, 128)
        self.conv4 = nn.Conv2d(128, c4, 3, 2, 1)
        self.norm4 = nn.GroupNorm(g, c4)
        self.res1_conv1 = nn.Conv2d(c4, c4, 3, 1, 1)
        self.res1_norm1 = nn.GroupNorm(g, c4)
        self.res1_conv2 = nn.Conv2d(c4, c4, 1, 1, 0)
        self.res1_norm2 = nn.GroupNorm(g, c4)
        self.res1_conv3 = nn.Conv2d(c4, c4, 3, 1, 1)
        self.res1_norm3 = nn.GroupNorm(g, c4)
        self.res2_conv1 = nn.Conv2d(c4, c5, 3, 1, 1)
        self.res2_norm1 = nn.GroupNorm(g, c5)
        self.res2_conv2 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.res2_norm2 = nn.GroupNorm(g, c5)
        self.res2_conv3 = nn.Conv2d(c5, c5, 3, 1, 1)
        self.res2_norm3 = nn.GroupNorm(g, c5)
        if not tiny:
            self.res2_skip = nn.Conv2d(256, 512, 1, 1, 0)
            self.res2_skip_norm = nn.GroupNorm(g, 512)
        self.enc_add_res_block_ls = [_res_block(tiny, g) for _ in range(enc_add_res_block)]
        for i, block in enumerate(self.enc_add_res_block_ls):
            self.add_module('enc_add_res_block{:d}'.format(i + 1), block)
        self._engine = None

    def plan(self, prefix):
        """(layers, blocks, roles) of this encoder for the native engine; names are state-dict keys under `prefix`."""
        pairs = [('conv1', 'norm1'), ('conv2', 'norm2'), ('conv3', 'norm3'), ('conv4', 'norm4'),
                 ('res1_conv1', 'res1_norm1'), ('res1_conv2', 'res1_norm2'), ('res1_conv3', 'res1_norm3'),
                 ('res2_conv1', 'res2_norm1'), ('res2_conv2', 'res2_norm2'), ('res2_conv3', 'res2_norm3')]
        layers = [(prefix + c, getattr(self, c), getattr(self, n)) for c, n in pairs]
        roles = {r: prefix + r for r in ('conv1', 'conv2', 'conv3', 'conv4')}
        res2 = {'kind': 'residual', 'convs': [prefix + 'res2_conv%d' % i for i in (1, 2, 3)]}
        if not self.tiny:
            layers.append((prefix + 'res2_skip', self.res2_skip, self.res2_skip_norm))
            res2 = {'kind': 'residual_skip', 'convs': res2['convs'], 'skip': prefix + 'res2_skip'}
        blocks = [{'kind': 'residual', 'convs': [prefix + 'res1_conv%d' % i for i in (1, 2, 3)]}, res2]
        for i, block in enumerate(self.enc_add_res_block_ls):
            names = [prefix + 'enc_add_res_block%d.%d' % (i + 1, j) for j in (0, 3, 6)]
            layers += [(names[k], block[3 * k], block[3 * k + 1]) for k in range(3)]
            blocks.append({'kind': 'residual', 'convs': names})
        return layers, blocks, roles

    def forward_reference(self, inputs, conv=_torch_conv):
        if inputs.size(0) == 0:
            raise RuntimeError('crossloc_b200: empty batch')
        x = F.relu(self.norm1(conv(self.conv1, inputs)))
        x = F.relu(self.norm2(conv(self.conv2, x)))
        x = F.relu(self.norm3(conv(self.conv3, x)))
        res = F.relu(self.norm4(conv(self.conv4, x)))
        x = F.relu(self.res1_norm1(conv(self.res1_conv1, res)))
        x = F.relu(self.res1_norm2(conv(self.res1_conv2, x)))
        x = F.relu(self.res1_norm3(conv(self.res1_conv3, x)))
        res = F.relu(res + x)
        x = F.relu(self.res2_norm1(conv(self.res2_conv1, res)))
        x = F.relu(self.res2_norm2(conv(self.res2_conv2, x)))
        x = F.relu(self.res2_norm3(conv(self.res2_conv3, x)))
        if not self.tiny:
            res = self.res2_skip_norm(conv(self.res2_skip, res))
        res = F.relu(res + x)
        for block in self.enc_add_res_block_ls:
            res = F.relu(res + _run_block(block, res, conv))
        return res

    def forward(self, inputs):
        """The encoder on its own (the reference calls it per MLR branch, networks.py:484-488): the native plan up to the
        residual stream, returned as an NCHW fp32 activation.  With autograd enabled: native convolutions per layer."""
        if not _native_ok(self, inputs):
            return self.forward_reference(inputs, conv=native_train.conv2d)
        if self._engine is None:
            self._engine = CoordNetEngine()
        layers, blocks, roles = self.plan('')
        return self._engine.forward({'group_norm': True, 'layers': layers, 'blocks': blocks, 'roles': roles,
                                     'output': 'activation'}, inputs)


class DenseUpsamplingConvolution(nn.Module):
    """DUC up-sampling head of the full-size variant (networks.py:259-273); semantics task only."""

    def __init__(self, down_sampling_rate, in_channel, num_classes, num_gn_channel=32):
        super(DenseUpsamplingConvolution, self).__init__()
        up = (down_sampling_rate ** 2) * num_classes
        self.conv = nn.Conv2d(in_channel, up, 3, 1, 1)
        self.norm = nn.GroupNorm(num_gn_channel, up)
        self.relu = nn.ReLU(inplace=True)
        self.pixel_shuffle = nn.PixelShuffle(down_sampling_rate)

    def forward(self, x, conv=_torch_conv):
        return self.pixel_shuffle(self.relu(self.norm(conv(self.conv, x))))


class TransPoseNetDecoder(nn.Module):
    """Decoder (networks.py:276-360): optional extra blocks, a 1x1 residual stage, fc1/fc2 and the output head."""

    def __init__(self, mean, tiny, dec_add_res_block=0, num_task_channel=3, num_pos_channel=1, num_gn_channel=32,
                 full_size_output=False):
        super(TransPoseNetDecoder, self).__init__()
        self.register_buffer('mean', mean.clone())
        self.tiny = tiny
        self.dec_add_res_block = dec_add_res_block
        self.num_task_channel = num_task_channel
        self.num_pos_channel = num_pos_channel
        self.num_gn_channel = num_gn_channel
        self.full_size_output = full_size_output
        c5, g = _width(tiny), num_gn_channel
        self.dec_add_res_block_ls = [_res_block(tiny, g) for _ in range(dec_add_res_block)]
        for i, block in enumerate(self.dec_add_res_block_ls):
            self.add_module('dec_add_res_block{:d}'.format(i + 1), block)
        self.res3_conv1 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.res3_norm1 = nn.GroupNorm(g, c5)
        self.res3_conv2 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.res3_norm2 = nn.GroupNorm(g, c5)
        self.res3_conv3 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.res3_norm3 = nn.GroupNorm(g, c5)
        self.fc1 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.fc1_norm = nn.GroupNorm(min(c5, g), c5)
        self.fc2 = nn.Conv2d(c5, c5, 1, 1, 0)
        self.fc2_norm = nn.GroupNorm(min(c5, g), c5)
        assert num_task_channel > 0 and num_pos_channel >= 0
        assert num_task_channel == len(mean)
        co = num_task_channel + num_pos_channel
        if full_size_output:
            self.duc_upsample = DenseUpsamplingConvolution(down_sampling_rate=8, in_channel=c5, num_classes=co)
            self.fc3 = nn.Conv2d(co, co, 1, 1, 0)
        else:
            self.fc3 = nn.Conv2d(c5, co, 1, 1, 0)
        self._engine = None

    def plan(self, prefix):
        layers, blocks = [], []
        for i, block in enumerate(self.dec_add_res_block_ls):
            names = [prefix + 'dec_add_res_block%d.%d' % (i + 1, j) for j in (0, 3, 6)]
            layers += [(names[k], block[3 * k], block[3 * k + 1]) for k in range(3)]
            blocks.append({'kind': 'residual', 'convs': names})
        names = [prefix + 'res3_conv%d' % i for i in (1, 2, 3)]
        layers += [(names[i], getattr(self, 'res3_conv%d' % (i + 1)), getattr(self, 'res3_norm%d' % (i + 1)))
                   for i in range(3)]
        blocks.append({'kind': 'residual', 'convs': names})
        layers += [(prefix + 'fc1', self.fc1, self.fc1_norm), (prefix + 'fc2', self.fc2, self.fc2_norm)]
        blocks.append({'kind': 'plain', 'convs': [prefix + 'fc1', prefix + 'fc2']})
        head = {'conv': self.fc3, 'mean': self.mean, 'num_task': self.num_task_channel, 'clamp': _POS_CLAMP}
        if self.full_size_output:
            name = prefix + 'duc_upsample.conv'
            layers.append((name, self.duc_upsample.conv, self.duc_upsample.norm))
            head['duc'] = {'name': name, 'rate': self.duc_upsample.pixel_shuffle.upscale_factor}
        return layers, blocks, head

    def forward_reference(self, inputs, up_height=None, up_width=None, conv=_torch_conv):
        res = inputs
        for block in self.dec_add_res_block_ls:
            res = F.relu(res + _run_block(block, res, conv))
        x = F.relu(self.res3_norm1(conv(self.res3_conv1, res)))
        x = F.relu(self.res3_norm2(conv(self.res3_conv2, x)))
        x = F.relu(self.res3_norm3(conv(self.res3_conv3, x)))
        res = F.relu(res + x)
        sc = F.relu(self.fc1_norm(conv(self.fc1, res)))
        sc = F.relu(self.fc2_norm(conv(self.fc2, sc)))
        if self.full_size_output:
            sc = self.duc_upsample(sc, conv)
            sc = F.interpolate(sc, (up_height, up_width), mode='bilinear', align_corners=False)
        sc = conv(self.fc3, sc)
        k = self.num_task_channel
        task = sc[:, :k] + self.mean.to(sc.device)[None, :, None, None]
        if not self.num_pos_channel:
            return task
        pos = torch.exp(F.hardtanh(sc[:, k:], min_val=_POS_CLAMP[0], max_val=_POS_CLAMP[1]))
        return torch.cat([task, pos], dim=1)

    def forward(self, inputs, up_height=None, up_width=None):
        """The decoder on its own (networks.py:319-360) on an NCHW activation: the native plan from the residual stream
        to the head.  With autograd enabled: native convolutions per layer."""
        if not _native_ok(self, inputs):
            return self.forward_reference(inputs, up_height, up_width, conv=native_train.conv2d)
        if self._engine is None:
            self._engine = CoordNetEngine()
        layers, blocks, head = self.plan('')
        if 'duc' in head:
            if up_height is None or up_width is None:
                raise RuntimeError('crossloc_b200: the full-size decoder needs up_height / up_width (networks.py:344-347)')
            head['duc']['size'] = (int(up_height), int(up_width))
        return self._engine.forward({'group_norm': True, 'layers': layers, 'blocks': blocks, 'head': head,
                                     'input': 'activation'}, inputs)


class TransPoseNet(nn.Module):
    """Encoder-decoder regression network with GroupNorm (networks.py:363-502)."""

    def __init__(self, mean, tiny, grayscale, enc_add_res_block=0, dec_add_res_block=0, num_task_channel=3,
                 num_pos_channel=1, num_gn_channel=32, num_mlr=0, num_unfrozen_encoder=0, full_size_output=False):
        super(TransPoseNet, self).__init__()
        self.register_buffer('mean', mean.clone())
        self.tiny = tiny
        self.grayscale = grayscale
        self.enc_add_res_block = enc_add_res_block
        self.dec_add_res_block = dec_add_res_block
        self.num_task_channel = num_task_channel
        self.num_pos_channel = num_pos_channel
        self.num_gn_channel = num_gn_channel
        self.num_mlr = num_mlr
        self.full_size_output = full_size_output
        self.OUTPUT_SUBSAMPLE = 1 if full_size_output else 8

        if num_mlr == 0:
            self.encoder = TransPoseNetEncoder(tiny, grayscale, enc_add_res_block, num_gn_channel)
        else:
            self.encoder = nn.Identity()
        self.encoder_ls = [self.encoder]

        if num_mlr > 0 and isinstance(num_mlr, int):
            assert 0 <= num_unfrozen_encoder <= num_mlr
            self.mlr_encoder_ls = [TransPoseNetEncoder(tiny, grayscale, enc_add_res_block, num_gn_channel)
                                   for _ in range(num_mlr)]
            for i, block in enumerate(self.mlr_encoder_ls):
                if i >= num_unfrozen_encoder:
                    for param in block.parameters():
                        param.requires_grad = False
                self.add_module('mlr_encoder_{:d}'.format(i + 1), block)
            width = _width(tiny)
            self.mlr_norm = nn.GroupNorm(num_gn_channel, width * num_mlr)
            self.mlr_forward = _res_block(tiny, num_gn_channel, in_ch=width * num_mlr)
            self.mlr_skip = nn.Sequential(nn.Conv2d(width * num_mlr, width, 1, 1, 0),
                                          nn.GroupNorm(num_gn_channel, width))
        else:
            self.mlr_encoder_ls = [nn.Identity()]
            self.mlr_norm = nn.Identity()
            self.mlr_forward = nn.Identity()
            self.mlr_skip = nn.Identity()
        self.mlr_ls = self.mlr_encoder_ls + [self.mlr_norm, self.mlr_forward, self.mlr_skip]

        self.decoder = TransPoseNetDecoder(mean, tiny, dec_add_res_block, num_task_channel, num_pos_channel,
                                           num_gn_channel, full_size_output)
        self.decoder_ls = [self.decoder]
        self._engine = None
        self._runtime = None

        count = sum(p.numel() for p in self.parameters() if p.requires_grad)
        safe_printout('Initialized TransPoseNet (crossloc_b200): tiny {}, grayscale {}, fullsize {}, #MLR {:d}, '
                      'extra blocks enc {:d} / dec {:d}, trainable parameters {:,d}.'.format(
                          tiny, grayscale, full_size_output, num_mlr, enc_add_res_block, dec_add_res_block, count))

    def _decoder_plan(self, inputs):
        layers, blocks, head = self.decoder.plan('decoder.')
        if 'duc' in head:   # the full-size head resizes to the frame size (networks.py:497-500)
            if inputs is None:
                raise RuntimeError('crossloc_b200: the full-size plan needs the input frame size')
            head['duc']['size'] = (int(inputs.size(2)), int(inputs.size(3)))
        return layers, blocks, head

    def _spec(self, inputs=None):
        enc_layers, enc_blocks, roles = self.encoder.plan('encoder.')
        dec_layers, dec_blocks, head = self._decoder_plan(inputs)
        return {'group_norm': True, 'layers': enc_layers + dec_layers, 'blocks': enc_blocks + dec_blocks, 'head': head,
                'roles': roles}

    def _forward_mlr(self, inputs):
        """MLR model (networks.py:482-494), fully on the fused plans: every encoder writes its output into a channel
        slice of one concatenated padded-flat activation (torch.cat, :488); the merge -- mlr_skip, mlr_norm over the
        concatenation (cl_pf_groupnorm), mlr_forward, relu(res + mlr) -- and the decoder run as one more plan on it."""
        eng = self._engine
        if eng.terms != 2:   # single-pass speed mode: no lo planes to normalise from
            return self._forward_mlr_unfused(inputs)
        width = _width(self.tiny)
        batch, _, h, w = inputs.shape
        concat = eng.shared_pf('mlr_concat', inputs.device, batch, h, w, width * self.num_mlr)
        skip_conv = self.mlr_skip[0]
        need8 = eng.nterms_of(skip_conv) == 2
        for i, enc in enumerate(self.mlr_encoder_ls):
            layers, blocks, roles = enc.plan('mlr_encoder_%d.' % (i + 1))
            eng.forward({'group_norm': True, 'layers': layers, 'blocks': blocks, 'roles': roles, 'output': 'pf',
                         'out_pf': (concat, i * width, need8)}, inputs)
        names = ['mlr_forward.%d' % j for j in (0, 3, 6)]
        merge_layers = [('mlr_skip.0', skip_conv, self.mlr_skip[1])]
        merge_layers += [(names[k], self.mlr_forward[3 * k], self.mlr_forward[3 * k + 1]) for k in range(3)]
        merge = {'kind': 'mlr_merge', 'skip': 'mlr_skip.0', 'convs': names, 'norm_in': self.mlr_norm}
        dec_layers, dec_blocks, head = self._decoder_plan(inputs)
        return eng.forward({'group_norm': True, 'layers': merge_layers + dec_layers, 'blocks': [merge] + dec_blocks,
                            'head': head, 'input': 'pf', 'in_pf': concat}, None)

    def _forward_mlr_unfused(self, inputs):
        """Encoders and decoder as fused plans, the merge on the native convolutions with stock GroupNorm."""
        acts = []
        for i, enc in enumerate(self.mlr_encoder_ls):
            layers, blocks, roles = enc.plan('mlr_encoder_%d.' % (i + 1))
            acts.append(self._engine.forward({'group_norm': True, 'layers': layers, 'blocks': blocks, 'roles': roles,
                                              'output': 'activation'}, inputs))
        mlr = torch.cat(acts, dim=1)
        conv = native_train.conv2d
        res = _run_block(self.mlr_skip, mlr, conv)
        mlr = _run_block(self.mlr_forward, self.mlr_norm(mlr), conv)
        res = F.relu(res + mlr)
        dec_layers, dec_blocks, head = self._decoder_plan(inputs)
        return self._engine.forward({'group_norm': True, 'layers': dec_layers, 'blocks': dec_blocks, 'head': head,
                                     'input': 'activation'}, res)

    def forward_reference(self, inputs, conv=_torch_conv):
        up_height, up_width = inputs.size()[2:4]
        if self.num_mlr == 0:
            res = self.encoder.forward_reference(inputs, conv)
        else:
            mlr = torch.cat([enc.forward_reference(inputs, conv) for enc in self.mlr_encoder_ls], dim=1)
            res = _run_block(self.mlr_skip, mlr, conv)
            mlr = _run_block(self.mlr_forward, self.mlr_norm(mlr), conv)
            res = F.relu(res + mlr)
        if self.full_size_output:
            return self.decoder.forward_reference(res, up_height, up_width, conv)
        return self.decoder.forward_reference(res, conv=conv)

    def forward_train(self, inputs, fused=None):
        """Autograd path on the native kernels.  Default: the fused plan (crossloc_b200.train_plan) -- the whole network
        is one autograd node whose backward runs GroupNorm / ReLU / residual-merge backward, data and weight gradients
        on padded-flat tensors; the 3-channel stem and the 4-channel head are differentiated with stock torch ops.
        fused=False (or the MLR / full-size variants): every eligible convolution as crossloc_b200.train.NativeConv2d
        between stock pointwise ops."""
        if fused is None:
            fused = _FUSED_TRAIN
        if fused and inputs.is_cuda and train_plan.supported(self):
            return train_plan.forward_train(self, inputs)
        return self.forward_reference(inputs, conv=native_train.conv2d)

    def forward(self, inputs):
        if not _native_ok(self, inputs):
            return self.forward_train(inputs)
        if self.num_mlr != 0:
            if self._engine is None:
                # fp16 + fp4 inside the encoders and the decoder; the merge convolutions that read the concatenated encoder
                # outputs / the mlr_norm result keep e4m3 corrections (those passes write no e2m1 planes)
                self._engine = CoordNetEngine(fp4=True)
                self._engine.no_fp4 = {id(self.mlr_skip[0]), id(self.mlr_forward[0])}
            return self._forward_mlr(inputs)
        return _run_native(self, self._spec(inputs), inputs)

    def forward_frames(self, frames_u8, mean=None, std=None):
        """uint8 HWC frames [B, H, W, C] (host or CUDA) -> network output: the host-to-device copy moves a quarter of the
        bytes and ToTensor [+ Normalize] run on the device (cl_net_forward_frames), bit-identical to torchvision."""
        if self.num_mlr != 0:
            raise RuntimeError('crossloc_b200: forward_frames covers the single-encoder network')
        if self._runtime is None:
            self._runtime = native_net.NetRuntime()
        b, h, w, _ = frames_u8.shape
        spec = self._spec(torch.empty(0, 3, h, w))
        return self._runtime.forward_frames(spec, frames_u8, mean, std)


class ProjHead(nn.Module):
    """Projection head (networks.py:505-541); unused by every reference script, kept for import compatibility."""

    def __init__(self, in_channel, out_length=2048, tiny=False, num_gn_channel=32):
        super(ProjHead, self).__init__()
        ch = _width(tiny)
        self.conv1 = nn.Conv2d(in_channel, ch, 3, 2, 1)
        self.norm1 = nn.GroupNorm(num_gn_channel, ch)
        self.conv2 = nn.Conv2d(ch, ch, 3, 2, 1)
        self.norm2 = nn.GroupNorm(num_gn_channel, ch)
        self.conv3 = nn.Conv2d(ch, ch, 3, 2, 1)
        self.norm3 = nn.GroupNorm(num_gn_channel, ch)
        self.conv4 = nn.Conv2d(ch, out_length, 1, 1, 0)
        self.norm4 = nn.GroupNorm(num_gn_channel, out_length)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.in_channel = in_channel
        self.out_length = out_length
        self.num_gn_channel = num_gn_channel

    def forward(self, inputs):
        x = F.relu(self.norm1(self.conv1(inputs)))
        x = F.relu(self.norm2(self.conv2(x)))
        x = F.relu(self.norm3(self.conv3(x)))
        x = F.relu(self.norm4(self.conv4(x)))
        return torch.flatten(self.avgpool(x), 1)

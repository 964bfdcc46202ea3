"""Public end-to-end API of the localization hot path: RGB frames in, camera poses out.

This is the body of the reference's evaluation loop (/root/reference/test_single_task.py:328-370:
`network(image.cuda())` -> `torch.split` -> `scene_coords_eval` -> `dsacstar.forward_rgb`) for a batch of
frames, with the coordinate map kept in HBM between the network and the solver (the reference's
`.cpu()` round trip, utils/evaluation.py:161, is gone).

`Localizer.localize(images_host)` takes pinned host images and returns host poses; `submit()` /
`result()` expose the same thing as a two-deep software pipeline so that the host-to-device copy of
the next batch overlaps the kernels of the current one.
"""
import os

import torch

from . import _lib, dsac


def frames_to_network_input(frames_u8, mean=None, std=None, out=None):
    """uint8 HWC frames [B,H,W,C] on the device -> fp32 NCHW network input: x / 255 (torchvision ToTensor) and, with
    mean / std, (v - mean) / std (Normalize) -- bit-identical to the host transform of dataloader/dataloader.py:189-212."""
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or not frames_u8.is_cuda:
        raise RuntimeError('frames must be a CUDA uint8 tensor [B, H, W, C]')
    frames_u8 = frames_u8.contiguous()
    b, h, w, c = frames_u8.shape
    if out is None:
        out = torch.empty(b, c, h, w, dtype=torch.float32, device=frames_u8.device)
    dev = frames_u8.device
    mean_t = None if mean is None else torch.as_tensor(mean, dtype=torch.float32, device=dev).contiguous()
    std_t = None if std is None else torch.as_tensor(std, dtype=torch.float32, device=dev).contiguous()
    _lib.check(_lib.load().cl_frames_to_nchw(frames_u8.data_ptr(), b, h, w, c, None if mean_t is None else mean_t.data_ptr(),
                                             None if std_t is None else std_t.data_ptr(), out.data_ptr(),
                                             torch.cuda.current_stream(dev).cuda_stream))
    return out


class Localizer:
    """Network + solver for batches of frames.

    Stream plan.  The CNN runs on the caller's stream; the pose solve (sample / score / refine: fp64 CUDA-core kernels,
    ~1 ms per 32 frames) runs on `solver_stream`.  With overlap=True the solve of batch k is DEFERRED: it is enqueued
    right after the CNN of batch k+1 and starts when that CNN enters its tensor-core-bound residual blocks
    (cl_net_wait_fork), where its blocks co-reside with the convolution CTAs (register / shared-memory budgets are
    chosen for that, csrc/dsac.cu) instead of delaying the stem of batch k+1.  `flush()` launches a solve that is still
    deferred (end of a sequence)."""

    def __init__(self, network, hyps=64, threshold=10.0, alpha=100.0, max_reproj=100.0, seed=1305, device=None,
                 defer_solve=True):
        self.net = network
        self.hyps, self.threshold, self.alpha, self.max_reproj, self.seed = hyps, threshold, alpha, max_reproj, seed
        self.device = torch.device(device if device is not None else 'cuda')
        self.subsample = int(getattr(network, 'OUTPUT_SUBSAMPLE', 8))
        self.num_task = int(getattr(network, 'num_task_channel', 3))
        self._copy_stream = torch.cuda.Stream(self.device)
        self.solver_stream = torch.cuda.Stream(self.device)
        self.solver_done = None   # event recorded after the last LAUNCHED solve; wait on it before reading poses
        self.defer_solve = defer_solve and os.environ.get('CROSSLOC_B200_DEFER_SOLVE', '1') != '0'
        self.after_solve = None   # optional callable(pose) run on the solver stream right behind every solve (pose gather)
        self._deferred = None
        self._slots = [None, None]
        self._turn = 0
        self._pending = []
        self.kernel_launches = 0
        self.solver_launches = 0

    # ------------------------------------------------------------------ solver scheduling
    def _launch_solve(self, job, wait_fork):
        coords, out_pose, focal, w, h, image_base, ready, done, extra = job
        with torch.cuda.stream(self.solver_stream):
            self.solver_stream.wait_event(ready)
            if wait_fork:
                rt = getattr(self.net, '_runtime', None)
                if rt is not None:
                    rt.wait_fork(self.solver_stream)
            for t in (coords, out_pose) + ((focal,) if torch.is_tensor(focal) and focal.is_cuda else ()):
                t.record_stream(self.solver_stream)
            dsac.forward_rgb_batch(coords, out_pose, self.hyps, self.threshold, focal, w / 2, h / 2, self.alpha,
                                   self.max_reproj, self.subsample, seed=self.seed, image_base=image_base)
            self.solver_launches += 3
            if self.after_solve is not None:
                self.after_solve(out_pose)
            if extra is not None:
                extra()
            done.record(self.solver_stream)
        self.solver_done = done

    def flush(self):
        """Launch the solve that is still deferred (if any); returns the event that marks the last solve's completion."""
        if self._deferred is not None:
            job, self._deferred = self._deferred, None
            self._launch_solve(job, wait_fork=False)
        return self.solver_done

    # ------------------------------------------------------------------ device-resident entry
    def localize_device(self, images, focal, coord_offset=None, image_base=0, out_pose=None, debug=False, overlap=False,
                        _extra=None, _done=None):
        """images [B,C,H,W] fp32 CUDA, focal float or [B]; returns poses [B,4,4] fp32 CUDA (camera-to-world).

        `coord_offset` ([B,3,Hc,Wc], optional) is added to the regressed coordinates before the solve: with
        random-initialised weights the raw map is geometrically meaningless, so the synthetic benchmark turns
        it into a consistent scene the same way the decoder's `mean` buffer offsets it.

        overlap=False: the poses are ready in the caller's stream on return (stream-ordered).
        overlap=True: the solve runs on `self.solver_stream`, possibly deferred until the next call; call `flush()`
        and wait on the returned event (or keep working on that stream) before touching the poses.
        """
        b, _, h, w = images.shape
        main = torch.cuda.current_stream(images.device)
        dbg = None
        with torch.no_grad():
            pred = self.net(images)
            self.kernel_launches = sum(e.launches for e in (getattr(self.net, '_engine', None), getattr(self.net, '_runtime', None))
                                       if e is not None)
            # the CNN of THIS batch is queued: the previous batch's deferred solve goes behind its fork point
            if self._deferred is not None:
                job, self._deferred = self._deferred, None
                self._launch_solve(job, wait_fork=True)
            coords = pred[:, :self.num_task]
            if coord_offset is not None:
                coords = coords + coord_offset
            coords = coords.contiguous()
            if out_pose is None:
                out_pose = torch.empty(b, 4, 4, dtype=torch.float32, device=images.device)
            ready = torch.cuda.Event()
            ready.record(main)
            done = _done if _done is not None else torch.cuda.Event()
            if debug:
                with torch.cuda.stream(self.solver_stream):
                    self.solver_stream.wait_event(ready)
                    for t in (coords, out_pose) + ((focal,) if torch.is_tensor(focal) and focal.is_cuda else ()):
                        t.record_stream(self.solver_stream)
                    dbg = dsac.forward_rgb_batch(coords, out_pose, self.hyps, self.threshold, focal, w / 2, h / 2,
                                                 self.alpha, self.max_reproj, self.subsample, seed=self.seed,
                                                 image_base=image_base, debug=True)
                    self.solver_launches += 3
                    done.record(self.solver_stream)
                self.solver_done = done
                main.wait_event(done)
                return out_pose, dbg
            job = (coords, out_pose, focal, w, h, image_base, ready, done, _extra)
            if overlap and self.defer_solve:
                self._deferred = job
            else:
                self._launch_solve(job, wait_fork=False)
                if not overlap:
                    main.wait_event(self.solver_done)
        return out_pose

    # ------------------------------------------------------------------ host entry, pipelined
    def submit(self, images_host, focal, coord_offset=None, image_base=0):
        """Queue one batch: async H2D on the copy stream, compute on the current stream, async D2H of the poses.
        `images_host`: pinned fp32 NCHW frames, or pinned uint8 HWC frames [B,H,W,C] (a quarter of the copy; converted on
        the device by cl_frames_to_nchw exactly as torchvision's ToTensor does on the host)."""
        slot = self._turn
        self._turn ^= 1
        b = images_host.size(0)
        st = self._slots[slot]
        if st is None or st['images'].shape != images_host.shape or st['images'].dtype != images_host.dtype:
            st = {
                'images': torch.empty(images_host.shape, dtype=images_host.dtype, device=self.device),
                'pose_dev': torch.empty(b, 4, 4, dtype=torch.float32, device=self.device),
                'pose_host': torch.empty(b, 4, 4, dtype=torch.float32).pin_memory(),
                'done': torch.cuda.Event(),
                'copied': torch.cuda.Event(),
                'free': None,
            }
            self._slots[slot] = st
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            if st['free'] is not None:
                self._copy_stream.wait_event(st['free'])   # previous use of this slot's device image has been consumed
            st['images'].copy_(images_host, non_blocking=True)
            st['copied'].record(self._copy_stream)
        compute.wait_event(st['copied'])
        frames = st['images']
        if frames.dtype == torch.uint8:
            st['frames_f32'] = frames = frames_to_network_input(frames, out=st.get('frames_f32'))

        def copy_back(st=st):   # poses leave on the solver stream, right behind their solve
            st['pose_host'].copy_(st['pose_dev'], non_blocking=True)

        self.localize_device(frames, focal, coord_offset, image_base, out_pose=st['pose_dev'], overlap=True,
                             _extra=copy_back, _done=st['done'])
        st['free'] = torch.cuda.Event()
        st['free'].record(compute)          # the network has consumed the device frames
        self._pending.append(st)
        return st

    def result(self):
        """Poses of the oldest submitted batch (host tensor, valid until its slot is reused two submits later)."""
        st = self._pending.pop(0)
        if self._deferred is not None and self._deferred[7] is st['done']:
            self.flush()   # nothing was submitted behind it: launch its solve now
        st['done'].synchronize()
        return st['pose_host']

    def localize(self, images_host, focal, coord_offset=None, image_base=0):
        """Synchronous form: one batch of pinned host images -> host poses."""
        self.submit(images_host, focal, coord_offset, image_base)
        return self.result()

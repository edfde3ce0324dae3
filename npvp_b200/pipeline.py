"""Lightning-free inference wrapper around the three hot-path modules.

``NPVPInference`` mirrors what ``LitPredictor`` does at inference time (models/Predictor.py:13-86):
it owns ``VPTR_Enc``, ``VPTR_Dec`` and ``predictor`` under the same attribute names (so a Lightning
checkpoint's ``state_dict`` loads after nothing more than ``torch.load(path)['state_dict']``), builds the
context / target timestamp lists from the YAML exactly like ``LitPredictor.__init__`` (:28-41) and exposes

  * ``forward(past, future=None)``  -> (rec_past, rec_future, pred_future)   the reference's triple (:72-86)
  * ``predict(past)``               -> predicted frames only, staying channels-last between the modules
  * ``rollout(past, num_future)``   -> block-autoregressive prediction beyond max_T (the shipped YAMLs ask for
                                       2 -> 28 frames with max_T 12)
Clips are independent, so multi-GPU inference shards the batch (``npvp_b200.distributed``).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from .autoencoder import ResnetDecoder, ResnetEncoder
from . import _lib
from .config import NORM, RENORM, AttrDict, load_config, preset
from .predictor import Predictor


def timestamp_lists(cfg):
    """(to_list, tp_list) as LitPredictor.__init__ builds them (Predictor.py:30-40)."""
    P, D = cfg.Predictor, cfg.Dataset
    if P.get("VFI", False):
        cp, cf, nv = P.context_num_p, P.context_num_f, P.num_interpolate
        clip = cp + cf + nv
        assert D.num_past_frames + D.num_future_frames == clip, "Imcompatible VFI configurations"
        idx = torch.linspace(0, clip - 1, clip, dtype=torch.int64)
        return torch.cat([idx[0:cp], idx[-cf:]]), idx[cp:-cf]
    to = torch.linspace(0, D.num_past_frames - 1, D.num_past_frames)
    tp = torch.linspace(D.num_past_frames, D.num_past_frames + D.num_future_frames - 1, D.num_future_frames)
    return to, tp


def _on_model_device(fn):
    """Run a method with the model's GPU as the current CUDA device (the C-ABI launches on the current device's stream)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapped


class NPVPInference(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        if isinstance(cfg, str):
            cfg = load_config(cfg) if cfg.endswith((".yaml", ".yml")) else preset(cfg)
        self.cfg = cfg
        A, D, P = cfg.AE, cfg.Dataset, cfg.Predictor
        self.VPTR_Enc = ResnetEncoder(D.img_channels, ngf=A.ngf, n_downsampling=A.n_downsampling,
                                      num_res_blocks=A.num_res_blocks, norm_layer=nn.BatchNorm2d,
                                      norm_layer1d=nn.BatchNorm1d, learn_3d=A.learn_3d)
        self.VPTR_Dec = ResnetDecoder(D.img_channels, ngf=A.ngf, n_downsampling=A.n_downsampling,
                                      out_layer=A.out_layer, norm_layer=nn.BatchNorm2d)
        self.h_list = torch.linspace(0, P.max_H - 1, P.max_H)
        self.w_list = torch.linspace(0, P.max_W - 1, P.max_W)
        self.to_list, self.tp_list = timestamp_lists(cfg)
        assert P.max_T == D.num_past_frames + D.num_future_frames, "Incompatible max_T and clip length"
        self.predictor = Predictor(P.max_H, P.max_W, P.max_T, self.h_list, self.w_list, self.to_list, self.tp_list,
                                   P.embed_dim, P.fuse_method, P.param_free_norm_type, P.evt_hidden_channels,
                                   1, P.stochastic, P.transformer_layers,
                                   evt_former=P.evt_former, learn_evt_token=False,
                                   evt_former_num_layers=P.evt_former_num_layers, rand_context=P.rand_context)
        if P.rand_context:
            self.predictor.reset_pos_coor(self.to_list, self.tp_list)
        # batch -> (context clip, target clip), selected like LitPredictor.__init__ (Predictor.py:62-70)
        if P.rand_context:
            self.batch_process_fn = self.rand_context_batch_process
        elif P.get("VFI", False):
            self.batch_process_fn = self.VFI_batch_process
        else:
            self.batch_process_fn = self.normal_batch_process
        self.eval()

    # -- batch pre-processing (Predictor.py:241-262) ----------------------------------------------------
    def rand_context_batch_process(self, batch):
        """(clip_o, clip_p, idx_o, idx_p) from the random-context dataloader: re-targets the predictor at the batch's context /
        target frame indices by slicing ``all_coor`` (Predictor.py:241-251) and returns (clip_o, clip_p)."""
        clip_batch_o, clip_batch_p, idx_o, idx_p = batch
        coor = self.predictor.all_coor
        self.predictor.observed_coor = coor[idx_o, ...].flatten(0, 2)
        self.predictor.predict_coor = coor[idx_p, ...].flatten(0, 2)
        self.predictor.TP = idx_p.shape[0]
        return (clip_batch_o, clip_batch_p)

    def VFI_batch_process(self, batch):
        """(past, future) -> (context frames, frames to interpolate) by the YAML's index lists (Predictor.py:253-259)."""
        past_frames, future_frames = batch
        clip = torch.cat([past_frames, future_frames], dim=1)
        return (clip[:, self.to_list, ...], clip[:, self.tp_list, ...])

    def normal_batch_process(self, batch):
        return batch

    # -- checkpoints --------------------------------------------------------------------------------
    def load_lightning_ckpt(self, path: str, strict: bool = True, trust_pickle: bool = False):
        """Load a reference Lightning ``.ckpt`` (keys VPTR_Enc.*, VPTR_Dec.*, predictor.*; Predictor.py:18-19,43).

        Only the ``state_dict`` entry is used; keys of other top-level modules of the LightningModule (losses, discriminator)
        are dropped.  The file is first read with ``weights_only=True`` (tensors and plain containers only).  Lightning
        checkpoints that pickle arbitrary objects next to the weights (hyper-parameter objects, callback state) need the
        unsafe unpickler like the reference's own ``load_from_checkpoint``: pass ``trust_pickle=True`` for files you trust."""
        import pickle
        try:
            blob = torch.load(path, map_location="cpu", weights_only=True)
        except (pickle.UnpicklingError, RuntimeError) as exc:
            if not trust_pickle:
                raise RuntimeError(f"{path}: not loadable with weights_only=True ({exc}); pass trust_pickle=True if the file is "
                                   "trusted (it is then unpickled like the reference does)") from exc
            blob = torch.load(path, map_location="cpu", weights_only=False)
        sd = blob.get("state_dict", blob) if isinstance(blob, dict) else blob
        keep = {k: v for k, v in sd.items() if k.split(".")[0] in ("VPTR_Enc", "VPTR_Dec", "predictor")}
        if not keep:
            raise KeyError(f"{path}: no VPTR_Enc.* / VPTR_Dec.* / predictor.* entries in the checkpoint's state_dict")
        return self.load_state_dict(keep, strict=strict)

    # -- reference-faithful forward -----------------------------------------------------------------
    @_on_model_device
    def forward(self, past_frames, future_frames=None):
        past_feats = self.VPTR_Enc(past_frames)
        rec_past = self.VPTR_Dec(past_feats)
        rec_future = None
        if future_frames is not None:
            rec_future = self.VPTR_Dec(self.VPTR_Enc(future_frames))
        pred = self.VPTR_Dec(self.predictor(past_feats))
        return rec_past, rec_future, pred

    # -- throughput path: Enc(context) -> Predictor -> Dec(predictions), channels-last in between ----
    def _norm_constants(self):
        """(mean, std) of VidNormalize for uint8 input frames (utils/dataset.py:34-58, 846-858)."""
        if self.cfg.AE.out_layer == "Sigmoid":
            return self._renorm_constants()
        return NORM.get(self.cfg.Dataset.name, RENORM[self.cfg.Dataset.name])

    def _predict_eager(self, past_frames, eps, pixels_u8=False):
        """``past_frames`` uint8: pixels, normalised inside the encoder stem.  ``pixels_u8``: the decoder head also writes the
        pixel-space uint8 frames -> (frames, frames_u8)."""
        self.predictor.injected_eps = eps
        try:
            self.predictor.prefetch_positional()
            norm = self._norm_constants() if past_frames.dtype == torch.uint8 else None
            feats = self.VPTR_Enc.forward_tokens(past_frames, norm=norm)
            pred = self.predictor.forward_tokens(feats, out16=self.VPTR_Dec._engine().dt)
            return self.VPTR_Dec.forward_tokens(pred, renorm=self._renorm_constants() if pixels_u8 else None)
        finally:
            self.predictor.injected_eps = None

    @_on_model_device
    def predict_samples(self, past_frames, n_samples: int, eps: Optional[torch.Tensor] = None):
        """NPVP-S: ``n_samples`` futures per clip -> (N, n_samples, Tp, Cimg, H, W).  The frame encoder, the EVT_Former and
        the prior run once per clip; latent sampling, the NAR decoder and the frame decoder run per sample (BASELINE
        config 3: 8 samples per clip).  ``eps``: optional (N*n_samples, 512, 8, 8) noise, clip-major."""
        self.predictor.injected_eps = eps
        try:
            self.predictor.prefetch_positional()
            feats = self.VPTR_Enc.forward_tokens(past_frames)
            pred = self.predictor.forward_tokens(feats, out16=self.VPTR_Dec._engine().dt, n_samples=n_samples)
            frames = self.VPTR_Dec.forward_tokens(pred)
            return frames.view(past_frames.shape[0], n_samples, *frames.shape[1:])
        finally:
            self.predictor.injected_eps = None

    MAX_GRAPHS = 8          # captured forwards kept per model (least recently used first out; each owns a private memory pool)

    @_on_model_device
    def predict_best_of_k(self, past_frames, future_frames, n_samples: int, metric: str = "psnr", eps: Optional[torch.Tensor] = None):
        """NPVP-S evaluation protocol for stochastic models (BASELINE config 3): draw ``n_samples`` futures per clip
        (``predict_samples``: frame encoder, EVT_Former and prior once per clip), score every sample against the ground-truth
        future frames in pixel space ([0,1], ``to_pixels``) with the reference's PSNR / SSIM (utils/metrics.py:12-109) and keep
        the best sample of every clip.  Returns (best frames in MODEL space (N, Tp, C, H, W), best_idx (N,), mean_scores (N, K))."""
        from .metrics import best_of_k
        smp = self.predict_samples(past_frames, n_samples, eps)
        if tuple(future_frames.shape) != tuple(smp.shape[:1] + smp.shape[2:]):
            raise ValueError(f"predict_best_of_k: future_frames must be {tuple(smp.shape[:1] + smp.shape[2:])}, got {tuple(future_frames.shape)}")
        _, idx, mean_scores, _ = best_of_k(self.to_pixels(smp), self.to_pixels(future_frames.to(smp.device)), metric)
        best = smp[torch.arange(smp.shape[0], device=smp.device), idx.long()]
        return best, idx, mean_scores

    def use_cuda_graphs(self, enabled: bool = True):
        """Replay ``predict`` as one CUDA graph per (batch shape, number of context / target timestamps, per-clip or shared
        timestamps, weights version): the ~400 kernel launches of a forward are launch-bound at small batch.  The returned
        tensor is then a graph-owned buffer that the next ``predict`` call with the same shapes overwrites.

        The timestamps themselves are NOT part of the key: the positional codes live in buffers owned by the captured forward
        and are recomputed in place, outside the graph, whenever ``reset_pos_coor`` / ``rand_context_batch_process`` installs
        new coordinates - a serving loop with random contexts (KTH unified) replays one graph.  At most ``MAX_GRAPHS`` captured
        forwards are kept; the least recently used one is dropped together with its memory pool."""
        self._graphs = {} if enabled else None
        return self

    def _graph_key(self, x, pixels_u8=False):
        p = self.predictor
        return (tuple(x.shape), x.dtype == torch.uint8, bool(pixels_u8), x.device, tuple(p.observed_coor.shape), tuple(p.predict_coor.shape), int(getattr(p, "_coor_clips", 0)),
                self.VPTR_Enc._weights_version(), self.VPTR_Dec._weights_version(), p._weights_version())

    @_on_model_device
    def predict(self, past_frames, eps: Optional[torch.Tensor] = None, pixels_u8: bool = False):
        """past_frames (N,To,Cimg,H,W) CUDA -> predicted frames (N,Tp,Cimg,H,W) fp32 (model space, like the reference modules).
        ``past_frames`` may be uint8 pixels (what a video decoder delivers): VidToTensor + VidNormalize then run inside the
        encoder's stem kernel.  ``pixels_u8=True``: returns (frames, frames_u8) where frames_u8 are the pixel-space uint8
        frames the reference would write to image files (VidReNormalize + clamp + ToPILImage, fused into the decoder head).
        ``eps``: optional injected latent noise (N,512,8,8) for NPVP-S (default: torch.randn like the reference)."""
        graphs = getattr(self, "_graphs", None)
        if graphs is None:
            return self._predict_eager(past_frames, eps, pixels_u8)
        self.predictor._coords_ready()
        key = self._graph_key(past_frames, pixels_u8)
        g = graphs.pop(key, None)
        if g is None:
            while len(graphs) >= self.MAX_GRAPHS:            # evict the least recently used captured forward (and its pool)
                graphs.pop(next(iter(graphs)))
            g = _GraphedPredict(self, past_frames, pixels_u8)
        graphs[key] = g                                       # (re-)insert as most recently used
        return g(past_frames, eps)

    def _short_block_coor(self, take: int):
        """Target coordinates of the first ``take`` timestamps the predictor is currently aimed at: the leading rows of its
        ``predict_coor`` (rows are timestamp-major), so the short block follows ``reset_pos_coor`` like the full blocks do.
        The slice shares the buffer's storage and modification counter, so its cached positional code is found again on the
        next rollout (engine_predictor.PredictorEngine.positional_table)."""
        p = self.predictor
        assert getattr(p, "_coor_clips", 0) == 0, "rollout(last_block='query') needs timestamps shared by the batch"
        rows = p.predict_coor.shape[0] // int(p.TP)
        return p.predict_coor[:take * rows]

    @_on_model_device
    def rollout(self, past_frames, num_future: int, eps_list: Optional[Sequence[torch.Tensor]] = None,
                out_host: Optional[torch.Tensor] = None, gather_group=None, last_block: str = "truncate",
                wait_output: bool = True, gather_dst: Optional[int] = 0, gather_dtype: torch.dtype = torch.float32):
        """Block-autoregressive VFP: predict len(tp_list) frames, feed the last To predictions back as context
        (image space), repeat until ``num_future`` frames exist.

        ``last_block`` - when fewer than len(tp_list) frames remain (2 -> 28 = 10 + 10 + 8 with the shipped YAMLs):
        "truncate" predicts the full block and drops the surplus frames; "query" asks the continuous-time decoder only for the
        timestamps that are needed (the first ``take`` of ``tp_list`` - SURVEY.md section 8d, "third block's tp cut to 8"), which
        skips the surplus frames' decoder and frame-decoder work.  The two differ in value (the target frames attend to each
        other over time), both are block-autoregressive uses of the reference model.

        ``past_frames`` may be a (pinned) host tensor: it is uploaded asynchronously.  With ``out_host`` (a pinned host
        tensor (N, num_future, C, H, W)) every block's frames are streamed to the host on a copy stream while the next block
        computes, and the caller's stream waits for the last copy before the call returns control of ``out_host``.
        A ``torch.uint8`` ``out_host`` receives pixel-space frames (``to_pixels(uint8=True)``, converted on the device block
        by block): a quarter of the D2H bytes of the fp32 model-space frames.

        ``wait_output=False`` (with ``out_host``): the caller's stream does not wait for the last block's device-to-host copy;
        ``self.output_ready`` (a CUDA event recorded on the copy stream) says when ``out_host`` is complete.  A serving loop
        that calls ``rollout`` back to back then overlaps the tail copy of one call with the first block of the next.

        ``gather_group`` (a ``torch.distributed`` process group, or ``True`` for the default group; every rank holds the
        same number of clips): the frames of all ranks are gathered BLOCK BY BLOCK with asynchronous collectives, so the
        exchange of block i overlaps the kernels of block i+1 (``distributed.BlockGather``).  ``gather_dst``: the group rank
        that receives the global batch (world * N, num_future, C, H, W), rank-major (default 0; the other ranks get their own
        shard back), or ``None``: every rank receives it (all-gather).  ``gather_dtype``: what travels - ``torch.float32``
        (model-space frames, bit-identical to the single-GPU result), ``torch.float16`` (model space, rounded) or
        ``torch.uint8`` (pixel-space frames like ``to_pixels(uint8=True)``: a quarter of the bytes)."""
        dev = next(self.parameters()).device
        if not past_frames.is_cuda:
            # upload on a stream of its own that waits for nothing: when calls are issued back to back (a serving loop) the copy of
            # this call's context overlaps the kernels of the previous call instead of sitting in front of the stem
            # (25 MB = ~1.2 ms of a 47 ms step).  ``self.input_consumed`` says when the host buffer may be overwritten.
            cur = torch.cuda.current_stream(dev)
            up = self.__dict__.setdefault("_upload_stream", torch.cuda.Stream(device=dev))
            with torch.cuda.stream(up):
                past_frames = past_frames.to(dev, non_blocking=True)
                self.input_consumed = torch.cuda.Event()
                self.input_consumed.record(up)
            past_frames.record_stream(cur)
            cur.wait_event(self.input_consumed)
        p = self.predictor
        p._coords_ready()
        if int(getattr(p, "_coor_clips", 0)):
            raise ValueError("rollout: block-autoregressive feedback needs timestamps shared by the batch (reset_pos_coor), not per-clip ones")
        To, Tp = int(p.observed_coor.shape[0]) // 64, int(p.TP)          # the predictor's CURRENT targeting (reset_pos_coor / batch fns)
        if past_frames.shape[1] != To:
            raise ValueError(f"rollout: the predictor is aimed at {To} context frames but the input has {past_frames.shape[1]}")
        coor_key = (p.observed_coor.data_ptr(), p.predict_coor.data_ptr(), int(p.observed_coor._version), int(p.predict_coor._version),
                    tuple(p.observed_coor.shape), tuple(p.predict_coor.shape))
        if self.__dict__.get("_rollout_checked") != coor_key:      # once per coordinate set: the check reads the device (a host sync)
            t_o, t_p = p.observed_coor[::64, 0].cpu(), p.predict_coor[::64, 0].cpu()
            if not (bool((t_o[1:] > t_o[:-1]).all()) and bool((t_p[1:] > t_p[:-1]).all()) and bool(t_p[0] > t_o[-1])):
                raise ValueError("rollout: feeding predictions back as context only makes sense for future prediction (increasing context "
                                 "timestamps followed by increasing target timestamps); the predictor is aimed at an interpolation / "
                                 "random-context task")
            self.__dict__["_rollout_checked"] = coor_key
            self.__dict__["_rollout_checked_refs"] = (p.observed_coor, p.predict_coor)   # keep the addresses from being recycled
        copy_stream = None
        if out_host is not None:
            assert not out_host.is_cuda and out_host.shape[1] == num_future
            assert out_host.dtype in (torch.float32, torch.uint8), "out_host must be fp32 (model space) or uint8 (pixel space)"
            copy_stream = self.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=dev))
        as_u8 = out_host is not None and out_host.dtype == torch.uint8
        out_u8 = None
        gatherer = None
        if gather_group is not None:
            import torch.distributed as dist
            from .distributed import BlockGather
            group = None if gather_group is True else gather_group
            assert gather_dtype in (torch.float32, torch.float16, torch.uint8), "gather_dtype: float32, float16 or uint8"
            if dist.get_world_size(group) > 1:
                gatherer = BlockGather(group, gather_dst, num_future)
        out, ctx, done, blk = None, past_frames, 0, 0
        assert last_block in ("truncate", "query")
        want_u8 = as_u8 or (gatherer is not None and gather_dtype == torch.uint8)
        while done < num_future:
            eps = None if eps_list is None else eps_list[blk]
            take = min(Tp, num_future - done)
            if take < Tp and last_block == "query":          # re-target the predictor at the timestamps still needed
                p = self.predictor
                saved = (p.predict_coor, p.TP)
                p.predict_coor, p.TP = self._short_block_coor(take), take
                try:
                    pred = self.predict(ctx, eps, pixels_u8=want_u8)
                finally:
                    p.predict_coor, p.TP = saved
            else:
                pred = self.predict(ctx, eps, pixels_u8=want_u8)   # may be graph-owned buffers: copied out before the next block
            if want_u8:                                      # pixel-space uint8 frames from the decoder head's fused epilogue
                pred, pred_u8 = pred
            if out is None:
                out = torch.empty((pred.shape[0], num_future) + tuple(pred.shape[2:]), dtype=pred.dtype, device=pred.device)
            out[:, done:done + take].copy_(pred[:, :take])
            src, s0 = out, done
            if want_u8:
                out_u8 = pred_u8.clone() if getattr(self, "_graphs", None) is not None else pred_u8   # graph-owned: the next block overwrites it
            if as_u8:
                src, s0 = out_u8, 0
            if copy_stream is not None:                      # D2H of this block overlaps the next block's kernels
                ready = torch.cuda.Event()
                ready.record()
                copy_stream.wait_event(ready)
                with torch.cuda.stream(copy_stream):
                    # per-clip copies: each (take, C, H, W) slab is contiguous on both sides, so every copy is one plain
                    # async memcpy (a strided (N, take, ...) slice would make torch stage through a temporary and synchronise)
                    for i in range(out.shape[0]):
                        out_host[i, done:done + take].copy_(src[i, s0:s0 + take], non_blocking=True)
                    if as_u8:
                        out_u8.record_stream(copy_stream)
            if gatherer is not None:
                # a private contiguous payload: `pred` may be a graph-owned buffer that the next block overwrites
                if gather_dtype == torch.uint8:
                    mine = out_u8[:, :take].contiguous() if take < out_u8.shape[1] else out_u8
                else:
                    mine = out[:, done:done + take].to(gather_dtype).contiguous() if gather_dtype != out.dtype else out[:, done:done + take].contiguous()
                gatherer.submit(mine, done)
            done += take
            blk += 1
            if done >= num_future:
                break
            if Tp >= To:
                ctx = out[:, done - To:done] if take == Tp else pred[:, Tp - To:Tp]
            else:
                ctx = torch.cat([ctx[:, Tp:], pred], dim=1)
        if copy_stream is not None:
            self.output_ready = torch.cuda.Event()
            self.output_ready.record(copy_stream)
            if wait_output:
                torch.cuda.current_stream().wait_stream(copy_stream)
            out.record_stream(copy_stream)
        if gatherer is not None:
            full = gatherer.result()
            if full is not None:
                return full
        return out

    # -- pixel space (utils/dataset.py:860-886, utils/train_summary.py:244-245) ----------------------
    def _renorm_constants(self):
        if self.cfg.AE.out_layer == "Sigmoid":
            c = int(self.cfg.Dataset.img_channels)
            return (0.0,) * c, (1.0,) * c
        return RENORM[self.cfg.Dataset.name]

    @_on_model_device
    def to_pixels(self, frames, uint8: bool = False):
        """Model output (..., C, H, W) -> [0,1] pixel space: VidReNormalize + clamp (utils/dataset.py:860-886,
        utils/train_summary.py:243-245; Sigmoid models: clamp only); ``uint8=True`` returns what the reference writes to
        image files (ToPILImage: trunc(v * 255)).  One CUDA kernel, reference operation order (bit-identical)."""
        if not frames.is_cuda:
            raise NotImplementedError("NPVPInference.to_pixels: frames must be a CUDA tensor (there is no CPU fallback)")
        mean, std = self._renorm_constants()
        x = frames.detach().to(torch.float32).contiguous()
        out = torch.empty_like(x, dtype=torch.uint8 if uint8 else torch.float32)
        _lib.ops().frames_to_pixels(x, mean, std, out_u8=out if uint8 else None, out_f32=None if uint8 else out)
        return out

    @_on_model_device
    def from_pixels(self, frames_u8):
        """uint8 frames (..., C, H, W) -> model-space fp32 input: VidToTensor + VidNormalize (utils/dataset.py:835-858)."""
        if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8:
            raise NotImplementedError("NPVPInference.from_pixels: expects a CUDA uint8 tensor")
        if self.cfg.AE.out_layer == "Sigmoid":
            mean, std = self._renorm_constants()
        else:
            mean, std = NORM.get(self.cfg.Dataset.name, RENORM[self.cfg.Dataset.name])
        x = frames_u8.contiguous()
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        _lib.ops().pixels_to_frames(x, mean, std, out)
        return out


class _GraphedPredict:
    """One captured forward: static input / noise buffers, graph-owned output, and graph-owned positional codes
    (``slots``): the NRMLP tables the captured kernels read are recomputed in place - eagerly, before the replay - when the
    predictor's coordinates have changed since the last call, so the graph does not depend on the timestamps."""

    def __init__(self, model: NPVPInference, example: torch.Tensor, pixels_u8: bool = False):
        self.model, self.pixels_u8 = model, bool(pixels_u8)
        self.stochastic = bool(model.predictor.stochastic)
        self.x = torch.empty_like(example, dtype=torch.uint8 if example.dtype == torch.uint8 else torch.float32).contiguous()
        self.x.copy_(example)
        n = example.shape[0]
        self.eps = torch.zeros((n, 512, 8, 8), device=example.device) if self.stochastic else None
        self.slots = ({}, {})
        self.eng = model.predictor._engine()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up: builds engines, workspaces, kernel attributes, positional codes
            for _ in range(2):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    def _run(self):
        self.eng._graph_slots = self.slots
        try:
            return self.model._predict_eager(self.x, self.eps, self.pixels_u8)
        finally:
            self.eng._graph_slots = None

    def __call__(self, x, eps):
        self.x.copy_(x)
        if self.stochastic:
            if eps is None:
                self.eps.normal_()                          # sampled outside the graph, like torch.randn in the reference
            else:
                self.eps.copy_(eps)
        p = self.model.predictor
        self.eng._graph_slots = self.slots
        try:                                                # no-op unless the coordinates changed: then 24 small eager launches
            self.eng._positional_pair(p.observed_coor, p.predict_coor)
        finally:
            self.eng._graph_slots = None
        self.graph.replay()
        return self.out


def build_from_config(cfg, device="cuda", seed: Optional[int] = 0) -> NPVPInference:
    """Random-init model of a config (YAML path, preset name or AttrDict), built in the order Enc, Dec, Predictor."""
    if seed is not None:
        torch.manual_seed(seed)
    return NPVPInference(cfg).to(device).eval()

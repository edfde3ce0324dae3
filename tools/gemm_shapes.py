"""List the distinct GEMM shapes (with epilogue flavour and call counts) of one predictor forward + autoencoder pass on
the CPU kernel specifications, scaled to a given clip count.  Used to decide where GEMM tuning pays."""
import collections
import sys
import os

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch

import npvp_b200._lib as _lib
from kernel_specs import SpecOps
from npvp_b200 import config
from npvp_b200.pipeline import NPVPInference

torch.set_grad_enabled(False)


class Rec(SpecOps):
    def __init__(self):
        super().__init__()
        self.shapes = collections.Counter()
        self.conv = False

    def gemm(self, a, w, *, bias=None, act=0, alpha=1.0, res1=None, res2=None, out_f32=None, out_bf16=None, post_relu=False, backend=None):
        M, K = a.shape
        N = w.shape[0]
        flav = f"act{act}" + ("+res" if res1 is not None else "") + ("+res2" if res2 is not None else "") + ("+relu" if post_relu else "") + \
               ("/f32" if out_f32 is not None else "/16")
        self.shapes[("conv" if self.conv else "dense", M, N, K, flav)] += 1
        return super().gemm(a, w, bias=bias, act=act, alpha=alpha, res1=res1, res2=res2, out_f32=out_f32, out_bf16=out_bf16,
                            post_relu=post_relu, backend=backend)

    def conv_gemm(self, *a, **k):
        self.conv = True
        try:
            return super().conv_gemm(*a, **k)
        finally:
            self.conv = False


    def convt_gemm(self, x, w, frames, H, W, Cin, Cout, **k):
        self.shapes[("convT 9/16", frames * H * W, 4 * Cout, 4 * Cin, f"act{k.get('act', 0)}/16")] += 1
        self.launches_before = self.launches
        return super().convt_gemm(x, w, frames, H, W, Cin, Cout, **k)


def main(preset="Cityscapes_VFP_NPVP-S", clips=1, scale=64):
    rec = Rec()
    _lib.set_ops(rec)
    cfg = config.preset(preset)
    m = NPVPInference(cfg).eval()
    D = cfg.Dataset
    x = torch.rand(clips, D.num_past_frames, D.img_channels, D.img_size, D.img_size)
    from npvp_b200.engine_autoencoder import DecoderEngine, EncoderEngine
    from npvp_b200.engine_predictor import PredictorEngine
    feats = EncoderEngine(m.VPTR_Enc).run(x)
    m.predictor.injected_eps = torch.randn(clips, 512, 8, 8)
    pred = PredictorEngine(m.predictor).run(feats)
    DecoderEngine(m.VPTR_Dec).run(pred)
    tot = 0.0
    rows = []
    for (kind, M, N, K, flav), c in rec.shapes.items():
        fl = 2.0 * M * scale * N * K * c
        tot += fl
        rows.append((fl, kind, M * scale, N, K, flav, c))
    print(f"| kind | M (x{scale} clips) | N | K | epilogue | calls | GFLOP | share |\n|---|---:|---:|---:|---|---:|---:|---:|")
    for fl, kind, M, N, K, flav, c in sorted(rows, reverse=True):
        print(f"| {kind} | {M} | {N} | {K} | {flav} | {c} | {fl / 1e9:.1f} | {100 * fl / tot:.1f}% |")
    print(f"total {tot / 1e12:.2f} TFLOP per forward at {scale} clips")


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["Cityscapes_VFP_NPVP-S"]))

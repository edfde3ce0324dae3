"""Golden fixtures for the pixel-space post-processing and metrics (tests/golden/post_*.npz), generated from the
UNMODIFIED reference (utils/metrics.py PSNR / SSIM, utils/dataset.py VidNormalize / VidReNormalize) and torchvision's
ToPILImage.  Build container only (needs /root/reference):  python tests/golden/make_golden_post.py
It also checks oracle/post_oracle.py against the reference on the same inputs (bit-exact for renorm / uint8 / normalise)."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("NPVP_REFERENCE", "/root/reference")

CASES = {   # name: (dataset constants as in utils/dataset.py:34-58, C, H, W, clips, T)
    "post_cityscapes": ("CityScapes", (0.31604213, 0.35114038, 0.3104223), (1.2172801, 1.3219808, 1.2082524), 3, 40, 36, 2, 3),
    "post_kth": ("KTH", (0.6013795,), (2.7570653,), 1, 33, 47, 2, 2),
    "post_smmnist": ("SMMNIST", (0.0,), (1.0,), 1, 24, 24, 1, 4),
}


def _load(name, path, package):
    spec = importlib.util.spec_from_file_location(f"{package}.{name}", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"{package}.{name}"] = mod
    spec.loader.exec_module(mod)
    return mod


def import_reference_utils():
    pkg = types.ModuleType("ref_utils")
    pkg.__path__ = []
    sys.modules["ref_utils"] = pkg
    ts = types.ModuleType("ref_utils.train_summary")       # metrics.py only needs the name at import time
    ts.load_ckpt = None
    sys.modules["ref_utils.train_summary"] = ts
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningDataModule = object
    pl.LightningModule = object
    sys.modules.setdefault("pytorch_lightning", pl)
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    metrics = _load("metrics", os.path.join(REF, "utils", "metrics.py"), "ref_utils")
    dataset = _load("dataset", os.path.join(REF, "utils", "dataset.py"), "ref_utils")
    return metrics, dataset


def main():
    import warnings
    warnings.filterwarnings("ignore")
    import torchvision.transforms as T
    from oracle import post_oracle as P
    metrics, dataset = import_reference_utils()
    lines = []
    for name, (ds, mean, std, C, H, W, N, Tt) in CASES.items():
        g = torch.Generator().manual_seed(7)
        frames = torch.randn(N, Tt, C, H, W, generator=g) * 0.6          # model-space "predictions"
        gt = torch.rand(N, Tt, C, H, W, generator=g)                     # pixel-space ground truth
        m_arg, s_arg = (mean, std) if C > 1 else (mean[0], std[0])
        renorm = dataset.VidReNormalize(mean=m_arg, std=s_arg)
        norm = dataset.VidNormalize(mean=m_arg, std=s_arg)
        pix = torch.stack([torch.clamp(renorm(frames[n].clone()), min=0., max=1.0) for n in range(N)])   # train_summary.py:243-245
        u8 = torch.stack([torch.stack([torch.from_numpy(np.array(T.ToPILImage()(pix[n, t]))).reshape(H, W, C).permute(2, 0, 1)
                                       for t in range(Tt)]) for n in range(N)])
        back = torch.stack([norm(torch.stack([T.ToTensor()(T.ToPILImage()(pix[n, t])) for t in range(Tt)])) for n in range(N)])
        flat_p, flat_g = pix.reshape(N * Tt, C, H, W), gt.reshape(N * Tt, C, H, W)
        ps = metrics.PSNR(flat_p, flat_g, mean_flag=False)
        ss = metrics.SSIM()(flat_p, flat_g, mean_flag=False)
        # oracle vs reference
        o_pix = P.renormalize_clamp(frames, mean, std)
        o_u8 = P.to_uint8(o_pix)
        o_back = P.normalize_u8(o_u8, mean, std)
        assert torch.equal(o_pix, pix), (name, float((o_pix - pix).abs().max()))
        assert torch.equal(o_u8, u8), name
        assert float((o_back - back).abs().max()) <= 1e-6, (name, float((o_back - back).abs().max()))
        o_ps, o_ss = P.psnr(flat_p, flat_g), P.ssim(flat_p, flat_g)
        assert float((o_ps - ps).abs().max()) <= 1e-5 and float((o_ss - ss).abs().max()) <= 1e-6, name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=frames.numpy(), gt=gt.numpy(), mean=np.array(mean, np.float64),
                            std=np.array(std, np.float64), pix=pix.numpy(), u8=u8.numpy(), back=back.numpy(), psnr=ps.numpy(),
                            ssim=ss.numpy())
        lines.append(f"{name}: dataset {ds} C={C} {H}x{W} frames={N * Tt}  oracle==reference: renorm/uint8 bit-exact, "
                     f"normalise {float((o_back - back).abs().max()):.1e}, psnr {float((o_ps - ps).abs().max()):.1e}, "
                     f"ssim {float((o_ss - ss).abs().max()):.1e}")
        print(lines[-1])
    with open(os.path.join(HERE, "REPORT_post.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

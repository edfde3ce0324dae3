"""CPU: host-side logic of the Lightning-free wrapper (no kernels are launched): batch pre-processing and coordinates."""
import torch

from npvp_b200.config import preset
from npvp_b200.pipeline import NPVPInference

torch.set_grad_enabled(False)


def test_rand_context_batch_process_retargets_coordinates():
    """LitPredictor.rand_context_batch_process (Predictor.py:241-251): coordinates follow the batch's frame indices."""
    m = NPVPInference(preset("KTH_Unified_NPVP-S"))
    assert m.batch_process_fn == m.rand_context_batch_process
    idx_o, idx_p = torch.tensor([0, 1, 2, 3, 6, 7, 10, 14, 15, 16]), torch.tensor([4, 5, 8, 9, 11, 12, 13, 17, 18, 19])
    clip = torch.zeros(2, 20, 1, 64, 64)
    o, p = m.batch_process_fn((clip[:, idx_o], clip[:, idx_p], idx_o, idx_p))
    assert o.shape[1] == 10 and p.shape[1] == 10 and m.predictor.TP == 10
    max_T = m.cfg.Predictor.max_T
    oc, pc = m.predictor.observed_coor, m.predictor.predict_coor
    assert oc.shape == (10 * 64, 3) and pc.shape == (10 * 64, 3)
    assert torch.allclose(oc[::64, 0], idx_o.float() / max_T) and torch.allclose(pc[::64, 0], idx_p.float() / max_T)
    # same coordinates as reset_pos_coor with the same timestamps (Predictor.py:352-359)
    m.predictor.reset_pos_coor(idx_o.float(), idx_p.float())
    assert torch.allclose(m.predictor.observed_coor, oc) and torch.allclose(m.predictor.predict_coor, pc)


def test_vfi_and_normal_batch_process():
    cfg = preset("KTH_Unified_NPVP-S")
    cfg.Predictor.rand_context = False
    cfg.Predictor.VFI = True
    cfg.Predictor.context_num_p, cfg.Predictor.context_num_f, cfg.Predictor.num_interpolate = 4, 4, 12
    m = NPVPInference(cfg)
    assert m.batch_process_fn == m.VFI_batch_process
    assert m.to_list.tolist() == [0, 1, 2, 3, 16, 17, 18, 19] and m.tp_list.tolist() == list(range(4, 16))
    past = torch.arange(10.).view(1, 10, 1, 1, 1).expand(1, 10, 1, 2, 2)
    fut = torch.arange(10., 20.).view(1, 10, 1, 1, 1).expand(1, 10, 1, 2, 2)
    o, p = m.batch_process_fn((past, fut))
    assert o[0, :, 0, 0, 0].tolist() == [0, 1, 2, 3, 16, 17, 18, 19] and p[0, :, 0, 0, 0].tolist() == list(range(4, 16))
    cfg.Predictor.VFI = False
    m2 = NPVPInference(cfg)
    assert m2.batch_process_fn((past, fut)) == (past, fut)


class _Opaque:
    """Stands in for the arbitrary objects (hyper-parameter namespaces, callback state) Lightning pickles into a .ckpt."""

    def __init__(self, v):
        self.v = v


def test_load_lightning_ckpt_round_trip(tmp_path):
    """A synthetic Lightning checkpoint {'state_dict': {VPTR_Enc.*, VPTR_Dec.*, predictor.*, <other modules>}, 'callbacks': ...,
    'optimizer_states': ...} (Predictor.py:18-19,43) loads with strict=True: extra top-level modules are dropped, the
    shared norm's two keys land in one tensor, the coordinate buffers of a checkpoint saved after reset_pos_coor (other
    shapes) are ignored, and a file that pickles arbitrary objects needs trust_pickle=True."""
    import pytest
    from util_init import stress_init_
    src = NPVPInference(preset("KITTI_VFP_NPVP-S"))
    for i, mod in enumerate((src.VPTR_Enc, src.VPTR_Dec, src.predictor)):
        stress_init_(mod, 50 + i)
    src.predictor.reset_pos_coor(torch.tensor([0., 1.]), torch.tensor([2., 3., 4.5]))       # buffers saved with other shapes
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    assert any(k.startswith("VPTR_Enc.") for k in sd) and "predictor.EVT_Former.norm.weight" in sd and "predictor.transformer.norm.weight" in sd
    sd["loss_fn.weight"] = torch.ones(3)                                                   # another module of the LightningModule
    sd["discriminator.model.0.weight"] = torch.zeros(4, 3, 4, 4)
    plain = {"epoch": 7, "global_step": 1234, "state_dict": sd, "optimizer_states": [{"state": {}, "param_groups": [{"lr": 1e-4}]}],
             "lr_schedulers": [], "callbacks": {"ModelCheckpoint": {"best_model_score": torch.tensor(0.5)}}}
    path = tmp_path / "plain.ckpt"
    torch.save(plain, path)
    dst = NPVPInference(preset("KITTI_VFP_NPVP-S"))
    shape_before = tuple(dst.predictor.predict_coor.shape)
    res = dst.load_lightning_ckpt(str(path))
    assert not res.missing_keys and not res.unexpected_keys
    ref = src.state_dict()
    for k, v in dst.state_dict().items():
        if k.endswith(("observed_coor", "predict_coor")):
            continue
        assert torch.equal(v, ref[k]), k
    assert tuple(dst.predictor.predict_coor.shape) == shape_before                           # derived buffers follow the module, not the file
    assert dst.predictor.EVT_Former.norm.weight.data_ptr() == dst.predictor.transformer.norm.weight.data_ptr()
    # a checkpoint with arbitrary pickled objects: refused by default, loaded with trust_pickle=True
    pickled = dict(plain, hyper_parameters=_Opaque(3))
    path2 = tmp_path / "pickled.ckpt"
    torch.save(pickled, path2)
    dst2 = NPVPInference(preset("KITTI_VFP_NPVP-S"))
    with pytest.raises(RuntimeError, match="trust_pickle"):
        dst2.load_lightning_ckpt(str(path2))
    dst2.load_lightning_ckpt(str(path2), trust_pickle=True)
    assert torch.equal(dst2.VPTR_Dec.state_dict()["model.0.weight"], ref["VPTR_Dec.model.0.weight"])
    # a bare module state_dict (no 'state_dict' wrapper) also loads; a file without the three prefixes is an error
    torch.save(src.state_dict(), tmp_path / "bare.ckpt")
    NPVPInference(preset("KITTI_VFP_NPVP-S")).load_lightning_ckpt(str(tmp_path / "bare.ckpt"))
    torch.save({"state_dict": {"foo.bar": torch.zeros(1)}}, tmp_path / "empty.ckpt")
    with pytest.raises(KeyError):
        NPVPInference(preset("KITTI_VFP_NPVP-S")).load_lightning_ckpt(str(tmp_path / "empty.ckpt"))

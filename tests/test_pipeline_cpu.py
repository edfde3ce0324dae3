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

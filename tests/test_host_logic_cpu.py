"""CPU: small pieces of host logic that the GPU tests only exercise indirectly."""
import pytest
import torch

from npvp_b200 import _lib
from npvp_b200.config import PRESETS, preset
from npvp_b200.pipeline import timestamp_lists
from npvp_b200.workspace import Workspace

torch.set_grad_enabled(False)


def test_workspace_keeps_every_shape_alive():
    """A captured CUDA graph replays with the raw pointers of the shapes it saw: asking for another shape under the same name
    must not free the first buffer (a rollout alternates between full and short blocks)."""
    ws = Workspace(torch.device("cpu"))
    a = ws.f32("y", 4, 8)
    b = ws.f32("y", 2, 8)
    assert a.data_ptr() != b.data_ptr()
    assert ws.f32("y", 4, 8) is a and ws.f32("y", 2, 8) is b            # stable pointers for both shapes
    assert ws.bf16("y", 4, 8).dtype == torch.bfloat16 and ws.bf16("y", 4, 8) is not a
    assert ws.bytes() == 4 * 8 * 4 + 2 * 8 * 4 + 4 * 8 * 2
    ws.trim()
    assert ws.bytes() == 0


def test_positional_code_period():
    """_pos_frames: the positional code covers T frames (timestamps shared by the batch) or n_clips * T (per-clip)."""
    T, n = 3, 4
    beta = torch.zeros(T * 64, 512)
    assert _lib._pos_frames(beta, None, n, T) == T
    assert _lib._pos_frames(torch.zeros(n * T * 64, 512), torch.zeros(n * T * 64, 512), n, T) == n * T
    assert _lib._pos_frames(None, None, n, T) == 0
    with pytest.raises(AssertionError):
        _lib._pos_frames(torch.zeros(2 * T * 64, 512), None, n, T)          # neither T nor n * T
    with pytest.raises(AssertionError):
        _lib._pos_frames(beta, torch.zeros(2 * T * 64, 512), n, T)          # gamma must match beta


@pytest.mark.parametrize("name", sorted(PRESETS))
def test_presets_are_consistent(name):
    """LitPredictor asserts max_T == num_past + num_future (Predictor.py:41); the timestamp lists follow Predictor.py:30-40."""
    cfg = preset(name)
    D, P = cfg.Dataset, cfg.Predictor
    assert P.max_T == D.num_past_frames + D.num_future_frames
    to, tp = timestamp_lists(cfg)
    assert to.tolist() == list(range(D.num_past_frames))
    assert tp.tolist() == list(range(D.num_past_frames, P.max_T))
    assert float(tp.max()) <= P.max_T


def test_per_clip_coordinates_layout():
    """reset_pos_coor_per_clip: clip-major concatenation of the per-clip coordinate tables; reset_pos_coor switches back."""
    import npvp_b200
    hl = torch.linspace(0, 7, 8)
    m = npvp_b200.Predictor(8, 8, 10, hl, hl, torch.tensor([0., 1.]), torch.tensor([2., 3., 4.]), 512, 'Add', 'layer', 256, 1, False, 1,
                            evt_former=True, learn_evt_token=False, evt_former_num_layers=1, rand_context=False)
    to, tp = torch.tensor([[0., 1.], [4., 9.]]), torch.tensor([[2., 3., 4.], [5., 6.5, 8.]])
    m.reset_pos_coor_per_clip(to, tp)
    assert m._coor_clips == 2 and m.TP == 3
    assert m.observed_coor.shape == (2 * 2 * 64, 3) and m.predict_coor.shape == (2 * 3 * 64, 3)
    assert torch.allclose(m.predict_coor[3 * 64::64, 0], tp[1] / 10) and torch.allclose(m.observed_coor[::64, 0], to.reshape(-1) / 10)
    with pytest.raises(AssertionError):
        m.reset_pos_coor_per_clip(torch.tensor([[0., 11.]]), tp[:1])              # outside max_T (submodules.py:351)
    m.reset_pos_coor(to[0], tp[0])
    assert m._coor_clips == 0 and m.predict_coor.shape == (3 * 64, 3)


def test_error_conventions_of_the_reference_api():
    """same exception types as the reference for bad arguments; unsupported reference options raise instead of
    silently computing something else."""
    import npvp_b200
    hl = torch.linspace(0, 7, 8)
    to, tp = torch.tensor([0., 1.]), torch.tensor([2., 3.])
    base = (8, 8, 4, hl, hl, to, tp, 512, 'Add')
    with pytest.raises(NotImplementedError):                    # ResNetAutoEncoder.py:239
        npvp_b200.ResnetEncoder(3, padding_type="circular", learn_3d=False)
    with pytest.raises(ValueError):                             # submodules.py:429
        npvp_b200.Predictor(*base, 'bogus', 256, 1, False, 1, evt_former_num_layers=1)
    with pytest.raises(NotImplementedError):
        npvp_b200.Predictor(*base, 'layer', 256, 1, False, 1, evt_former=False)
    with pytest.raises(NotImplementedError):
        npvp_b200.Predictor(*base, 'layer', 256, 1, False, 1, evt_former_num_layers=1, learn_evt_token=True)
    with pytest.raises(NotImplementedError):
        npvp_b200.Predictor(*base, 'layer', 256, 1, False, 1, evt_former_num_layers=1, return_intermediate=True)
    with pytest.raises(NotImplementedError):
        npvp_b200.Predictor(8, 8, 4, hl, hl, to, tp, 256, 'Add', 'layer', 256, 1, False, 1, evt_former_num_layers=1)   # embed_dim != 512
    rc = npvp_b200.Predictor(*base, 'layer', 256, 1, True, 1, evt_former_num_layers=1, rand_context=True).eval()
    assert rc.observed_coor is None and rc.predict_coor is None and tuple(rc.all_coor.shape) == (4, 8, 8, 3)   # Predictor.py:282-284
    with pytest.raises(RuntimeError, match="reset_pos_coor"):   # the reference would crash inside nrmlp(None)
        rc._coords_ready()
    rc.reset_pos_coor(to, tp)
    assert rc.observed_coor.shape == (2 * 64, 3) and rc.TP == 2
    rc.train()
    with pytest.raises(NotImplementedError, match="inference"):  # training-mode posterior sampling (Predictor.py:316-318) is rejected
        rc(torch.zeros(1, 2, 512, 8, 8), torch.zeros(1, 2, 512, 8, 8))

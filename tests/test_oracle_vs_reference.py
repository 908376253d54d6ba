"""Pins the oracle against the UNMODIFIED reference, executed in-process.  Runs only where the
reference is mounted (/root/reference, i.e. the build container); skipped on the GPU box, where the
golden vectors made by tests/golden/make_golden.py stand in for it."""
import pytest
import torch

from oracle import ref_harness as rh
from oracle import spair_oracle as so

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference not mounted")

TINY = dict(INPUT_IMAGE_SHAPE=[1, 40, 40], OBJECT_SHAPE=[8, 8], ANCHORBOX_SHAPE=[16, 16],
            DEFAULT_BACKBONE_TOPOLOGY=rh.TOPOLOGY_CELL8, BATCH_SIZE=3)


@pytest.mark.parametrize("step", [1, 1001, 4000])
def test_oracle_equals_reference_forward_and_backward(step):
    ns = rh.load_reference(TINY)
    net = rh.build_reference_model(ns, seed=3)
    cfg = so.config_tiny()
    x = so.scattered_sprites(3, cfg.image_shape, seed=5, sprite_px=(6, 14))
    loss, recon, z_where, z_pres = rh.run_reference(net, x, step, noise_seed=11)
    ref_grads = {k: (None if p.grad is None else p.grad.clone()) for k, p in net.named_parameters()}
    params = so.params_from_state_dict(net.state_dict())
    out = so.forward_backward(params, x, step, so.draw_noise(11, 3, cfg.grid, cfg.n_attr), cfg)
    assert torch.equal(out["loss"], loss) and torch.equal(out["recon_x"], recon)
    assert torch.equal(out["z_where"], z_where) and torch.equal(out["z_pres"], z_pres)
    for n, d in net.dist.items():                       # the reference's own distribution maps (models.py:122-125)
        assert torch.equal(d.loc, out["dist_mean"][n]) and torch.equal(d.scale, out["dist_std"][n])
    for k, g in ref_grads.items():
        if g is None:
            assert params[k].grad is None
        else:
            assert float((g - params[k].grad).abs().max()) <= 2e-6 * float(g.abs().max()) + 1e-12, k


def test_oracle_stn_equals_reference_stn():
    ns = rh.load_reference(dict(INPUT_IMAGE_SHAPE=[3, 64, 64]))
    g = torch.Generator().manual_seed(0)
    img = torch.rand(5, 3, 64, 64, generator=g)
    zw = torch.rand(5, 4, generator=g) * torch.tensor([1.0, 1.0, 0.5, 0.5]) + torch.tensor([0.0, 0.0, 0.05, 0.05])
    dev = torch.device("cpu")
    a = ns.modules.stn(img, zw, [14, 14], dev)
    assert torch.equal(a, so.stn(img, zw, [14, 14]))
    assert torch.equal(ns.modules.stn(a, zw, [64, 64], dev, inverse=True), so.stn(a, zw, [64, 64], inverse=True))


def test_seeded_construction_reproduces_reference_parameters():
    """Same torch.manual_seed(3) -> the drop-in model has bit-identical parameters (train.py:39-41)."""
    from tests import helpers
    ns = rh.load_reference(TINY)
    ref = rh.build_reference_model(ns, seed=3)
    ours = helpers.build_model("tiny")
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs) == list(os_)
    for k in rs:
        assert torch.equal(rs[k], os_[k]), k


def test_metrics_match_reference():
    """spair.metric (evaluation path, reference metric.py:5-100) on random boxes."""
    from spair_pytorch_b200 import config as cfg
    from spair_pytorch_b200 import metric as ours
    ns = rh.load_reference(dict(BATCH_SIZE=4))
    g = torch.Generator().manual_seed(1)
    saved = cfg.BATCH_SIZE
    cfg.BATCH_SIZE = 4
    try:
        z_where = torch.rand(4, 4, 11, 11, generator=g) * 0.4
        z_pres = torch.rand(4, 1, 11, 11, generator=g)
        gt = torch.rand(4, 9, 4, generator=g) * 40 + 5
        count = torch.randint(1, 10, (4, 1), generator=g).float()
        a = ns.metric.mAP(z_where.clone(), z_pres.clone(), gt.clone(), count.clone())
        b = ours.mAP(z_where.clone(), z_pres.clone(), gt.clone(), count.clone())
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
        assert torch.equal(ns.metric.object_count_accuracy(z_pres, count), ours.object_count_accuracy(z_pres, count))
        box_a, box_b = torch.rand(4, 7, 4, generator=g), torch.rand(4, 5, 4, generator=g)
        box_a[..., 2:] += box_a[..., :2]
        box_b[..., 2:] += box_b[..., :2]
        assert torch.allclose(ns.metric.batch_jaccard(box_a, box_b), ours.batch_jaccard(box_a, box_b), rtol=1e-6, atol=1e-7)
    finally:
        cfg.BATCH_SIZE = saved


def test_checkpoint_saved_by_the_reference_loads_into_the_drop_in(tmp_path):
    """reference train.py:85-90 saves ``spair_net.state_dict()``; the drop-in model must load that file strictly (same
    keys, shapes, dtypes) and the other way round."""
    from tests import helpers
    ns = rh.load_reference({})
    ref = rh.build_reference_model(ns, seed=11)
    path = str(tmp_path / "step_1000.pkl")
    torch.save(ref.state_dict(), path)
    ours = helpers.build_model("A", seed=3)
    before = ours.box_network.body[0].weight.detach().clone()
    missing, unexpected = ours.load_state_dict(torch.load(path), strict=True)
    assert not missing and not unexpected
    assert not torch.equal(before, ours.box_network.body[0].weight)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
    ref2 = rh.build_reference_model(rh.load_reference({}), seed=5)
    ref2.load_state_dict(ours.state_dict(), strict=True)
    for k, v in ref2.state_dict().items():
        assert torch.equal(v, ref.state_dict()[k]), k

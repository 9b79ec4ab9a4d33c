"""SURVEY 8(a) a15/a16/a18 pinned to the reference's OWN loop: `tests/golden/trajectory_small.npz` holds what the unmodified
`engine.train_one_epoch` + `NativeScalerWithGradNormCount` + `lr_sched.adjust_learning_rate` + AdamW(0.9, 0.95) produced when
driving the reference model for 2 epochs x 8 micro-batches (accum_iter 2) on the golden config (oracle/make_golden.py).

  * CPU: `flipped_vqa_b200.engine.train_one_epoch` + our scaler / LR schedule drive the ORACLE (fp32, same arithmetic as the
    golden run) -> every per-step loss, every logged statistic and the final trainables must coincide (1e-5): this isolates
    the HOST loop (accumulation boundaries, LR timing, zero_grad, update order).
  * GPU: the same loop drives the product model -> losses within north_star's 1e-2; the trainables move the same way.
"""
import argparse
import os

import numpy as np
import pytest
import torch

from oracle import llama_vqa_oracle as O      # checker only
from oracle.make_golden import TRAJ, trajectory_args, trajectory_batches
from tests.util_parity import GOLDEN, GOLDEN_DIR, GOLDEN_RUN, LOSS_RTOL, build_product_model, golden_inputs, make_args, rel_l2


def _gold():
    return np.load(os.path.join(GOLDEN_DIR, "trajectory_small.npz"))


def _adamw(trainables):
    # `train.py:120-121`: timm's weight-decay grouping puts every trainable of this model (all >= 2-D) in the decayed group
    return torch.optim.AdamW([{"params": trainables, "weight_decay": TRAJ["weight_decay"]}], lr=TRAJ["lr"], betas=(0.9, 0.95))


class _OracleModel:
    """Duck-typed stand-in for the model: the oracle's functional step over a state dict of leaf tensors."""

    def __init__(self, sd, params):
        self.st, self.params, self.step_losses = O.prepare_state(sd), params, []
        self.names = O.trainable_names(self.st)

    def train(self, mode=True):
        return self

    def parameters(self):
        return [self.st[n] for n in self.names]

    def __call__(self, data):
        r = GOLDEN_RUN
        out = O.forward_losses(self.st, self.params, data, max_feats=r["max_feats"], tau=r["tau"])
        self.step_losses.append([float(x.detach()) for x in out])
        return out


def _run(model, trainables, record):
    from flipped_vqa_b200 import engine
    from flipped_vqa_b200.util import misc
    opt = _adamw(trainables)
    scaler = misc.NativeScalerWithGradNormCount()
    batches, targs = trajectory_batches(), trajectory_args()
    stats = []
    for epoch in range(TRAJ["epochs"]):
        stats.append(engine.train_one_epoch(model, batches, opt, epoch, scaler, args=targs))
        record(epoch)
    return stats


def test_engine_loop_reproduces_reference_trajectory_on_cpu():
    g = _gold()
    params, sd, _ = golden_inputs()
    model = _OracleModel(sd, params)
    snap = {}
    stats = _run(model, model.parameters(), lambda e: snap.update({(e, n): model.st[n].detach().clone() for n in model.names}))
    got = np.array(model.step_losses)
    assert got.shape == g["step_losses"].shape == (TRAJ["epochs"] * TRAJ["n_batches"], 3)
    assert np.abs(got - g["step_losses"]).max() / np.abs(g["step_losses"]).max() < 1e-5
    for e, st in enumerate(stats):
        keys = {k.split("/")[-1] for k in g.files if k.startswith(f"epoch{e}/stats/")}
        assert set(st) == keys == {"lr", "loss", "vqa_loss", "vaq_loss", "qav_loss"}
        for k in keys:
            assert abs(st[k] - float(g[f"epoch{e}/stats/{k}"])) <= 1e-5 * max(1.0, abs(st[k])), (e, k)
        for n in model.names:
            key = f"epoch{e}/param/{n}"
            if key in g.files:
                assert rel_l2(snap[(e, n)], g[key]) < 1e-5, (e, n)
    # layers skipped by `model.py:338` never move
    assert torch.equal(model.st["layers.0.attention.gate1"], O.prepare_state(sd)["layers.0.attention.gate1"])


@pytest.mark.gpu
def test_product_training_trajectory_matches_reference(fvqa_lib):
    g = _gold()
    params, sd, _ = golden_inputs()
    r = GOLDEN_RUN
    model = build_product_model(GOLDEN, sd, make_args(r["max_feats"], r["bias"], r["tau"]))
    losses = []
    fwd = model.forward

    def recording_forward(data, inference=False):
        out = fwd(data, inference)
        losses.append(out)
        return out
    model.forward = recording_forward
    init = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    snap = {}
    stats = _run(model, [p for p in model.parameters() if p.requires_grad],
                 lambda e: snap.update({(e, n): p.detach().clone().cpu() for n, p in model.named_parameters() if p.requires_grad}))
    got = np.array([[float(x) for x in o] for o in losses])
    ref = g["step_losses"]
    assert got.shape == ref.shape
    assert (np.abs(got - ref) / np.abs(ref)).max() < LOSS_RTOL, (np.abs(got - ref) / np.abs(ref)).max()
    for e, st in enumerate(stats):
        for k in ("loss", "vqa_loss", "vaq_loss", "qav_loss", "lr"):
            assert abs(st[k] - float(g[f"epoch{e}/stats/{k}"])) <= LOSS_RTOL * abs(float(g[f"epoch{e}/stats/{k}"])), (e, k)
    # The trainables follow the reference's path. AdamW's first updates are sign-like (lr * g / |g|), so elements whose gradient
    # is within rounding of zero may step the other way: the bound is on the direction and size of the total displacement.
    e = TRAJ["epochs"] - 1
    for n, p0 in init.items():
        key = f"epoch{e}/param/{n}"
        if key not in g.files:
            continue
        ref_p = torch.from_numpy(g[key])
        d_ref, d_got = (ref_p - p0.cpu()).flatten(), (snap[(e, n)] - p0.cpu()).flatten()
        if float(d_ref.norm()) == 0.0:
            assert float(d_got.norm()) == 0.0, n
            continue
        cos = float(torch.dot(d_ref, d_got) / (d_ref.norm() * d_got.norm()))
        assert cos > 0.98, (n, cos)
        assert abs(float(d_got.norm() / d_ref.norm()) - 1.0) < 0.05, n
        assert rel_l2(snap[(e, n)], ref_p) < 1e-2, n


@pytest.mark.parametrize("reference_style_scaler", [False, True])
def test_non_finite_loss_exits_before_the_optimizer_is_touched(reference_style_scaler):
    """`engine.py:33-35`: 'Loss is nan, stopping training' + sys.exit(1), and the non-finite loss never reaches the trainables or the
    AdamW state - with our scaler (check between backward and the update) and with any scaler that only has the reference's
    signature (check before backward, like the reference)."""
    from flipped_vqa_b200 import engine
    from flipped_vqa_b200.util import misc

    class _NaNOnSecondStep(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(4))
            self.calls = 0

        def forward(self, data):
            self.calls += 1
            base = (self.w * self.w).sum()
            bad = base * float("nan") if self.calls == 2 else base
            return bad, base.detach() * 0 + 1.0, base.detach() * 0 + 2.0

    model = _NaNOnSecondStep()
    opt = torch.optim.AdamW(model.parameters(), lr=0.1)
    ours = misc.NativeScalerWithGradNormCount()
    steps = []

    def ref_style(loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):   # `util/misc.py:259-273`
        steps.append("scaler")
        return ours(loss, optimizer, clip_grad=clip_grad, parameters=parameters, create_graph=create_graph, update_grad=update_grad)

    args = argparse.Namespace(accum_iter=1, lr=0.1, min_lr=0.0, warmup_epochs=0, epochs=1, debug=False)
    with pytest.raises(SystemExit) as ex:
        engine.train_one_epoch(model, [{}, {}, {}], opt, 0, ref_style if reference_style_scaler else ours, args=args)
    assert ex.value.code == 1
    assert model.calls == 2
    assert torch.isfinite(model.w).all()                              # the NaN step never reached the parameters ...
    st = opt.state[model.w]
    assert int(st["step"]) == 1 and torch.isfinite(st["exp_avg"]).all() and torch.isfinite(st["exp_avg_sq"]).all()   # ... nor AdamW's moments
    if reference_style_scaler:
        assert steps == ["scaler"]                                    # second call never happened: the check ran before backward

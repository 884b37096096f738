"""The reference's own unit-test factory, run on the mirror CLASSES through the C ABI.

Port of ``make_test_case_bijector`` (unittest_flow_models.py:25-51) and of its cases (:122-186): the toy coupling network
returning constants ``log_s = log 2, t = 1`` is injected through the ``shift_and_log_scale_layer`` factory seam
(:76-83, flow_tfp_bijectors.py:132), minibatches of 2s and 1s make ActNorm's scale exactly 2 (:66-73).

Two departures from the letter of the reference file, both documented in SURVEY.md section 4 ("staleness"): the
multi-scale classes are given the ``n_filters`` argument their constructors now require (flow_glow.py:82,147), and
the equality checks carry an fp32 tolerance -- ``inverse(forward(x)) == x`` bit-for-bit does not hold for
``(x * 2 - 3 + 3) / 2`` in any float32 arithmetic.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOG2 = float(np.log(2, dtype=np.float32))
EVENT_SHAPE, EVENT_SHAPE_1, EVENT_SHAPE_2, EVENT_SHAPE_3 = [2, 2, 1], [4, 4, 1], [2, 2, 2], [8, 8, 1]


def _inputs(shape, seed):
    return torch.randn([1] + shape, generator=torch.Generator().manual_seed(seed))


def _minibatch(shape):
    return torch.cat((2 * torch.ones([1] + shape), torch.ones([1] + shape)), dim=0)


def shift_and_log_scale_toy(x):
    return LOG2 * torch.ones(x.shape), torch.ones(x.shape)


def shift_and_log_scale_layer_toy(event_shape, n_hidden_units=2, name="toy", l2_reg=None, **kwargs):
    return shift_and_log_scale_toy


def make_test_case_bijector(bijector_class, inputs, expected_log_det, **kwargs):
    """unittest_flow_models.py:25-51 as a pair of checks."""
    bijector = bijector_class(**kwargs)

    # the multi-scale toy cases are badly conditioned by construction: ActNorm is initialised from a TWO-sample minibatch
    # (std = half the distance of two numbers) and every toy coupling doubles the state, so fp32 round-off of the
    # forward pass is amplified by the inverse; a single bijector inverts to fp32 round-off
    deep = bijector_class.__name__.startswith("GlowBijector")

    def check_inversibility():
        outputs = bijector.forward(inputs)
        inv_outputs = bijector.inverse(outputs)
        np.testing.assert_allclose(inv_outputs.cpu().numpy(), inputs.cpu().numpy(), rtol=5e-3 if deep else 1e-5,
                                   atol=1e-3 if deep else 2e-6)
        assert not torch.equal(outputs.cpu().reshape(-1), inputs.reshape(-1))       # the bijector did something

    def check_log_det():
        fldj = bijector.forward_log_det_jacobian(inputs, event_ndims=3).cpu().numpy()[0]
        outputs = bijector.forward(inputs)
        ildj = bijector.inverse_log_det_jacobian(outputs, event_ndims=3).cpu().numpy()[0]
        assert fldj == pytest.approx(-ildj, rel=1e-5, abs=1e-5)
        if expected_log_det is not None:
            assert fldj == pytest.approx(expected_log_det, rel=1e-5, abs=1e-5)

    return check_inversibility, check_log_det, bijector


def _cases():
    from audiosourcesep_b200.flow_models.flow_glow import GlowBijector_2blocks, GlowBijector_3blocks, GlowBlock, GlowStep
    from audiosourcesep_b200.flow_models.flow_tfp_bijectors import (ActNorm, AffineCouplingLayerSplit, Invertible1x1Conv,
                                                                    SpecPreprocessing, Squeeze)
    toy = dict(shift_and_log_scale_layer=shift_and_log_scale_layer_toy, n_hidden_units=2)
    return {
        # :141-146 one coupling layer on (2,2,2): 4 variables scaled by 2 -> 4 log 2
        "AffineCouplingLayerSplit": (AffineCouplingLayerSplit, _inputs(EVENT_SHAPE_2, 0), 4 * LOG2,
                                     dict(event_shape=EVENT_SHAPE_2, **toy)),
        # :149-154 ActNorm initialised to scale 2 on (2,2,1): 4 log 2
        "ActNorm": (ActNorm, _inputs(EVENT_SHAPE, 1), 4 * LOG2, dict(event_shape=EVENT_SHAPE, minibatch=_minibatch(EVENT_SHAPE))),
        # :157-161 invertibility (|det W| = 1 for the QR initialisation: the log-det is 0 up to round-off)
        "Invertible1x1Conv": (Invertible1x1Conv, _inputs(EVENT_SHAPE_2, 2), None, dict(event_shape=EVENT_SHAPE_2)),
        # :164-168 GlowStep on (2,2,2): actnorm 4*2*log 2 + 1x1 ~0 + coupling 4 log 2 = 12 log 2 (closed form, not in the
        # reference file, follows from the two cases above)
        "GlowStep": (GlowStep, _inputs(EVENT_SHAPE_2, 3), None, dict(event_shape=EVENT_SHAPE_2, minibatch=_minibatch(EVENT_SHAPE_2), **toy)),
        "GlowBlock": (GlowBlock, _inputs(EVENT_SHAPE_1, 4), None, dict(K=2, event_shape=EVENT_SHAPE_1, minibatch=_minibatch(EVENT_SHAPE_1), **toy)),
        "GlowBijector_2blocks": (GlowBijector_2blocks, _inputs(EVENT_SHAPE_1, 5), None,
                                 dict(K=2, event_shape=EVENT_SHAPE_1, n_filters=2, minibatch=_minibatch(EVENT_SHAPE_1), **toy)),
        "GlowBijector_3blocks": (GlowBijector_3blocks, _inputs(EVENT_SHAPE_3, 6), None,
                                 dict(K=2, event_shape=EVENT_SHAPE_3, n_filters=2, minibatch=_minibatch(EVENT_SHAPE_3), **toy)),
        "Squeeze": (Squeeze, _inputs(EVENT_SHAPE_1, 7), 0.0, dict(event_shape_in=EVENT_SHAPE_1)),
        "SpecPreprocessing": (SpecPreprocessing, 120.0 * torch.rand([1] + EVENT_SHAPE_1) - 100.0, 16 * math.log(1.0 / 120.0),
                              dict(minval=-100.0, maxval=20.0, use_logit=False)),
    }


CASE_NAMES = ["AffineCouplingLayerSplit", "ActNorm", "Invertible1x1Conv", "GlowStep", "GlowBlock", "GlowBijector_2blocks",
              "GlowBijector_3blocks", "Squeeze", "SpecPreprocessing"]


@pytest.mark.parametrize("name", CASE_NAMES)
def test_bijector_class(name):
    cls, inputs, expected, kwargs = _cases()[name]
    inv, logdet, _ = make_test_case_bijector(cls, inputs, expected, **kwargs)
    inv()
    logdet()


def test_glow_step_closed_form_log_det_and_order():
    """GlowStep = Chain([coupling, inv1x1, actnorm]) (flow_glow.py:21-22): forward runs actnorm -> 1x1 -> coupling, and
    with the toy network the log-det is 8 log 2 (ActNorm, scale 2 on 2x2x2) + 0 (orthogonal 1x1) + 4 log 2."""
    from audiosourcesep_b200.flow_models.flow_glow import GlowStep
    step = GlowStep(event_shape=EVENT_SHAPE_2, shift_and_log_scale_layer=shift_and_log_scale_layer_toy,
                    minibatch=_minibatch(EVENT_SHAPE_2), n_hidden_units=2)
    x = _inputs(EVENT_SHAPE_2, 11)
    fldj = step.forward_log_det_jacobian(x, event_ndims=3).cpu().numpy()[0]
    assert fldj == pytest.approx(12 * LOG2, abs=2e-5)
    manual = step.coupling_layer.forward(step.inv1x1conv.forward(step.actnorm.forward(x.cuda())))
    assert torch.equal(manual, step.forward(x.cuda()))


def test_glow_block_runs_steps_in_reverse_list_order():
    """GlowBlock = Chain(glow_steps + [squeeze]) (flow_glow.py:51-52): forward = squeeze, then step K-1 ... step 0."""
    from audiosourcesep_b200.flow_models.flow_glow import GlowBlock
    blk = GlowBlock(K=3, event_shape=EVENT_SHAPE_1, shift_and_log_scale_layer=shift_and_log_scale_layer_toy,
                    minibatch=_minibatch(EVENT_SHAPE_1), n_hidden_units=2)
    x = _inputs(EVENT_SHAPE_1, 12).cuda()
    h = blk.squeeze.forward(x)
    for k in (2, 1, 0):
        h = blk.glow_steps[k].forward(h)
    assert torch.equal(h, blk.forward(x))
    assert tuple(h.shape) == (1, 2, 2, 4)


@pytest.mark.parametrize("L,shape", [(2, EVENT_SHAPE_1), (3, EVENT_SHAPE_3)])
def test_composable_multiscale_matches_fused_handle_layout(L, shape):
    """The composable multi-scale classes factor out with a plain row-major reshape (flow_glow.py:179,182) and
    concatenate z1 | z2 | z3 on channels: output shape [N, H/2^L, W/2^L, C 4^L] and exact invertibility of the
    split / reshape plumbing, block by block."""
    from audiosourcesep_b200.flow_models import flow_glow
    cls = {2: flow_glow.GlowBijector_2blocks, 3: flow_glow.GlowBijector_3blocks}[L]
    bij = cls(K=1, event_shape=shape, shift_and_log_scale_layer=shift_and_log_scale_layer_toy, n_filters=2,
              minibatch=_minibatch(shape), n_hidden_units=2)
    x = _inputs(shape, 13)
    z = bij.forward(x)
    s = 1 << L
    assert tuple(z.shape) == (1, shape[0] // s, shape[1] // s, shape[2] * s * s)
    # first block by hand: z1 is the first half of its channels, reshaped row-major
    o1 = bij.blocks[0].forward(x.cuda())
    C1 = o1.shape[-1]
    z1 = o1[..., : C1 // 2].reshape(1, shape[0] // s, shape[1] // s, -1)
    assert torch.equal(z[..., : z1.shape[-1]], z1)
    np.testing.assert_allclose(bij.inverse(z).cpu().numpy(), x.numpy(), rtol=1e-5, atol=2e-6)


def test_fused_glow_bijector_class_round_trip_and_log_det():
    """GlowBijector_3blocks with the real coupling network (the fused libasep handle, flow_glow.py:145-225): TFP
    protocol on the class, fldj = -ildj, round trip, ActNorm initialised from the minibatch."""
    from audiosourcesep_b200 import synthetic
    from audiosourcesep_b200.flow_models.flow_glow import GlowBijector_3blocks
    from audiosourcesep_b200.flow_models.flow_tfk_layers import ShiftAndLogScaleConvNet
    shape = [16, 8, 1]
    mb = torch.as_tensor(synthetic.normalise(synthetic.mel_patches_db(8, seed=5, H=16, W=8))) - 0.5
    bij = GlowBijector_3blocks(K=2, event_shape=shape, shift_and_log_scale_layer=ShiftAndLogScaleConvNet, n_filters=64,
                               minibatch=mb)
    x = torch.as_tensor(synthetic.normalise(synthetic.mel_patches_db(3, seed=6, H=16, W=8))) - 0.5
    z = bij.forward(x)
    assert tuple(z.shape) == (3, 2, 1, 64)
    np.testing.assert_allclose(bij.inverse(z).cpu().numpy(), x.numpy(), atol=1e-4)
    fldj = bij.forward_log_det_jacobian(x, event_ndims=3).cpu().numpy()
    ildj = bij.inverse_log_det_jacobian(z, event_ndims=3).cpu().numpy()
    np.testing.assert_allclose(fldj, -ildj, rtol=1e-4, atol=1e-3)

"""Host-side mirror of the hot-path part of ``flow_models/flow_tfk_layers.py``.

The reference module defines the coupling network ``ShiftAndLogScaleConvNet`` (flow_tfk_layers.py:31-84: conv 3x3 ->
ReLU -> BatchNorm -> conv 1x1 -> ReLU -> BatchNorm -> conv 3x3, tanh on the log-scale) next to ResNet / Flow++ layers
that no melspec config uses.  Here the network is ONE fused tcgen05 kernel per Glow step
(csrc/nn_tc.cu, csrc/nn_tcx.cu) owned by the Glow handle, so the class is the token the reference's call sites pass as
``shift_and_log_scale_layer`` (flow_builder.py:118-125, flow_glow.py:84-86): ``GlowBijector_*blocks(K, event_shape,
ShiftAndLogScaleConvNet, n_filters, minibatch)`` selects the fused path.  A stand-alone evaluation of one step's
network is ``Glow.coupling_nn(block, step, state)`` (C ABI: ``asep_glow_coupling_nn``).
"""
from .flow_glow import ShiftAndLogScaleConvNet

__all__ = ["ShiftAndLogScaleConvNet"]

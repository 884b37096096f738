from . import flow_builder, flow_glow, flow_tfp_bijectors  # noqa: F401

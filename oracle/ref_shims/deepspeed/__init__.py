"""Import shim: the reference imports deepspeed at module scope (infer_utils.py:6) but never uses it at inference."""


def add_config_arguments(parser):
    return parser

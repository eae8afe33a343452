from .mrla_light_module import mrla_light_layer  # noqa: F401

from .mrla_light_module import mrla_light_layer  # noqa: F401
from .mrla_base_module import mrla_base_layer  # noqa: F401

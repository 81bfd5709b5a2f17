"""Mirror of the reference package ``nr4seg.nerf`` (network, renderer, raymarching, activation)."""
from .network_tcnn_semantics import SemanticNeRFNetwork  # noqa: F401
from .renderer_semantics import SemanticNeRFRenderer  # noqa: F401

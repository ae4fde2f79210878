"""Everything the reference imports from diffusers.models.embeddings is vendored in architecture/embeddings.py."""
from architecture.embeddings import (  # noqa: F401
    PixArtAlphaTextProjection,
    TimestepEmbedding,
    Timesteps,
    get_1d_rotary_pos_embed,
)

from .agg_block import AggregationBlock  # noqa: F401
from .attention import Attention, FeedForward, PostNorm, PreNorm, cache_fn  # noqa: F401
from .pos_encoding import build_position_encoding  # noqa: F401

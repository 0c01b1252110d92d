"""Position encodings of the aggregation block (reference: agg_block/pos_encoding.py:127-138).

DEVIAS always builds the block with pos_enc_type='none' (agg_block/agg_block.py:21,53), for which the
reference returns `lambda x: None`; that behaviour is kept.  The sine / learned variants are never
enabled by any run script and are rejected explicitly instead of silently
diverging from the reference."""


def build_position_encoding(dim, pos_type, axis):
    if pos_type in ('none',):
        return lambda x: None
    if pos_type in ('sine', 'learned'):
        raise NotImplementedError(
            f"pos_enc_type={pos_type!r} is outside the DEVIAS hot path (only 'none' is used by the reference recipes)")
    raise ValueError(f"not supported {pos_type}")

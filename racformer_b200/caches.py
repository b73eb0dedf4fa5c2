"""Epoch of the weight-derived caches of this package (bf16x3 operand splits, transposes, folded biases).

Every such cache is keyed on (data_ptr, autograd version counter, ..., cache_epoch()). The version counter misses updates made
through `.data` (`p.data.copy_()` / `p.data.mul_()` in EMA or weight-averaging utilities); `invalidate_weight_caches()` covers
those. decoder.RaCFormerTransformer calls it from `load_state_dict` and `train()`; a captured graphs.GraphedDecoderForward
bakes the old operand pointers into its graph and refuses to replay after the epoch has changed (re-capture it)."""
_EPOCH = [0]


def cache_epoch():
    return _EPOCH[0]


def invalidate_weight_caches():
    _EPOCH[0] += 1

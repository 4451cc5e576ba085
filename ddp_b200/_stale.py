"""Staleness check of a cached CUDA engine against the nn.Module parameters it was built from.

The engines copy the weights into the library's own layouts, so they must be rebuilt whenever a parameter changes.
``load_state_dict`` / ``_apply`` hooks do not see every path: mmcv's ``load_checkpoint`` recurses over
``_load_from_state_dict`` without calling the model's ``load_state_dict`` (mmcv/runner/checkpoint.py), and in-place edits
(``p.data.copy_``, ``p.mul_``) bypass both.  Every such path bumps the tensor's version counter or moves its storage, so
the (data_ptr, _version) list of the parameters is a cheap, sufficient fingerprint (≈150 tensors, tens of microseconds).
"""


def fingerprint(module, skip_prefixes=()):
    out = []
    for name, t in list(module.named_parameters()) + list(module.named_buffers()):
        if skip_prefixes and name.startswith(skip_prefixes):
            continue
        out.append((name, t.data_ptr(), t._version))
    return tuple(out)

"""Host-side string helper of the eval tail (reference: ``utils/util.py:12-24`` ``str_filt``) — pure Python in the reference as well."""
from __future__ import annotations

import string

_VOC = {'digit': string.digits, 'lower': string.digits + string.ascii_lowercase, 'upper': string.digits + string.ascii_letters,
        'all': string.digits + string.ascii_letters + string.punctuation}


def str_filt(str_: str, voc_type: str) -> str:
    """Keep the characters of ``voc_type``'s vocabulary ('lower' lower-cases first); unknown vocabularies raise KeyError like the reference."""
    keep = _VOC[voc_type]
    if voc_type == 'lower':
        str_ = str_.lower()
    return ''.join(c for c in str_ if c in keep)

"""Constructor-side helpers with the reference's semantics
(/root/reference/models_pytorch/utils/tools.py:5-13)."""


def pair(val):
    """int -> (int, int); tuples pass through (tools.py:5-6)."""
    return val if isinstance(val, tuple) else (val, val)


def check_sizes(image_size, patch_size):
    """Number of patches; asserts divisibility exactly like the reference (tools.py:8-13)."""
    ih, iw = pair(image_size)
    ph, pw = pair(patch_size)
    assert ih % ph == 0 and iw % pw == 0, 'image height and width must be divisible by patch size'
    return (ih // ph) * (iw // pw)

"""dataset.data_utils stand-in (the real module imports pythonOCC at its top, ref: dataset/data_utils.py:7-12)."""


def parse_splits_list(splits):
    """Here a 'split' is a range of synthetic drawing indices: 'a:b'."""
    a, b = str(splits).split(':')
    return list(range(int(a), int(b)))

from pamrec_b200.sequential_iterator import SequentialIterator, lisan, bar_border_list, takatak_bar_border_list_dict  # noqa: F401


class SASequentialIterator(SequentialIterator):
    """Imported by the driver but only used by the SASREC baseline (stale in the reference: io/sequential_iterator.py:1467)."""

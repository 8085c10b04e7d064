from pamrec_b200.models import ShareBottomModel  # noqa: F401

from pamrec_b200.models import PLEModel  # noqa: F401

from pamrec_b200.models import SASRecModel  # noqa: F401

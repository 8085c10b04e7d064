from pamrec_b200.models import PAMRECModel  # noqa: F401

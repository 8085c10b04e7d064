from pamrec_b200.models import MMoEModel_original  # noqa: F401

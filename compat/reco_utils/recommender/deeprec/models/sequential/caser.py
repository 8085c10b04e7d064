class CaserModel:
    """Baseline model of the reference, outside the PAMRec hot path."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("CaserModel is out of scope of pamrec_b200; only PAMRECModel is implemented")

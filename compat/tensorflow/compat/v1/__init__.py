from pamrec_b200.models import latest_checkpoint as _latest

__version__ = "none (pamrec_b200 shim)"


def disable_v2_behavior():
    return None


class train:  # noqa: N801  (tf.train.latest_checkpoint)
    latest_checkpoint = staticmethod(_latest)

from pamrec_b200.deeprec_utils import *  # noqa: F401,F403
from pamrec_b200.deeprec_utils import prepare_hparams, load_dict, cal_metric, cal_weighted_metric  # noqa: F401

SEED = 42   # reco_utils/common/constants.py:23 in the reference (imported by the driver, unused)

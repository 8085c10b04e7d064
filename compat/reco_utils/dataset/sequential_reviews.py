def data_preprocessing(*args, **kwargs):
    """The reference's preprocessing only knows taobao / kuaishou (dataset/sequential_reviews.py:48-51) and cannot build
    the wechat / takatak files either; use pamrec_b200.synth.generate for synthetic data in the same formats."""
    raise NotImplementedError("train_data is missing: generate it with pamrec_b200.synth.generate(root, dataset)")

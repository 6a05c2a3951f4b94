"""Mirror of ``nasrec/utils/config.py``: embedding-table cardinalities.

The reference caps table sizes through a source constant that users edit by hand
(config.py:17-19: ``500000 * 10000`` as shipped = uncapped; ``500000`` for search).
Here the cap is the environment variable ``NASREC_MAX_NUM_EMBEDDINGS`` (default:
uncapped, as shipped) or ``capped(...)``.
"""
import os

MAX_NUM_EMBEDDINGS = int(os.environ.get("NASREC_MAX_NUM_EMBEDDINGS", 500000 * 10000))

_CRITEO = [1461, 584, 10131227, 2202609, 306, 25, 12518, 634, 4, 93146, 5684, 8351593, 3195, 28, 14993, 5461307, 11,
           5653, 2174, 5, 7046548, 19, 16, 286182, 106, 142573]                       # config.py:21-23
_AVAZU = [10000, 241, 8, 8, 4738, 7746, 27, 8553, 560, 37, 2686409, 6729487, 8252, 6, 5, 2627, 9, 10, 436, 5, 69, 173,
          61]                                                                          # config.py:30-31
_KDD = [26274, 641708, 14848, 22122011, 1188090, 3735797, 2934102, 20004011, 4, 8]     # config.py:37


def capped(sizes, cap=None):
    cap = MAX_NUM_EMBEDDINGS if cap is None else cap
    return [min(x, cap) for x in sizes]


NUM_EMBEDDINGS_CRITEO = capped(_CRITEO)
NUM_EMBEDDINGS_AVAZU = capped(_AVAZU)
NUM_EMBEDDINGS_KDD = capped(_KDD)
NUM_EMBEDDINGS_TEST = [100] * 26                                                       # config.py:41

# entry scripts hard-code these (train_supernet.py:226-230, main_train.py:223-227)
NUM_SPARSE_INPUTS = {"criteo-kaggle": 26, "avazu": 23, "kdd": 10}
NUM_DENSE_INPUTS = {"criteo-kaggle": 13, "avazu": 1, "kdd": 3}
NUM_EMBEDDINGS = {"criteo-kaggle": _CRITEO, "avazu": _AVAZU, "kdd": _KDD}

"""Mirror of ``nasrec/supernet/utils.py``: any-path cardinality samplers (same calls
into numpy's global legacy RNG) and the ops-config validator."""
import numpy as np


def _get_random_choice_vanilla(num_items, max_items=4):
    """utils.py:21-27: uniform over 1..min(num_items, max_items)."""
    return np.random.choice(min(num_items, max_items)) + 1


def _get_binomial_random_choice_with_expectation(num_items, p=0.5, max_items=4):
    """utils.py:30-35: 1 + Binomial(min(num_items, max_items) - 1, p)."""
    return 1 + np.random.binomial(min(num_items - 1, max_items - 1), p)


anypath_choice_fn = {                                   # utils.py:38-43
    "uniform": lambda num_items: _get_random_choice_vanilla(num_items, max_items=4),
    "binomial-0.5": lambda num_items: _get_binomial_random_choice_with_expectation(num_items, p=0.5, max_items=4),
}


def assert_valid_ops_config(ops_config):                # utils.py:46-61
    for key, cfg in ops_config.items():
        for c in (cfg if isinstance(cfg, list) else [cfg]):
            assert c["num_nodes"] == len(c["node_names"]), ValueError(
                "Number of nodes per config should be equivalent to the number of modules (node names) per config.")

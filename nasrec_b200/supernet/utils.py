"""Mirror of ``nasrec/supernet/utils.py``: any-path cardinality samplers (same calls
into numpy's global legacy RNG) and the ops-config validator."""
import numpy as np


# numpy's legacy global RNG draws, call for call what the reference's np.random.choice(...) consumes -- the
# legacy RandomState implements choice(n) as randint(0, n), choice(n, k) as randint(0, n, size=k) and
# choice(n, k, replace=False) as permutation(n)[:k] -- without choice()'s ~6 us of argument handling per call
# (a training step draws ~60 numbers; tests/test_host_parity.py pins the streams against the reference's).
def pick_index(n: int) -> int:
    """== int(np.random.choice(n))"""
    return int(np.random.randint(0, n))


def pick(seq):
    """== np.random.choice(seq) for a list"""
    return seq[int(np.random.randint(0, len(seq)))]


def pick_with_replacement(n: int, k: int) -> np.ndarray:
    """== np.random.choice(n, k)"""
    return np.random.randint(0, n, size=k)


def pick_without_replacement(pop, k: int) -> list:
    """== np.random.choice(pop, k, replace=False).tolist() for an int or a list population"""
    if isinstance(pop, (int, np.integer)):
        return np.random.permutation(int(pop))[:k].tolist()
    idx = np.random.permutation(len(pop))[:k]
    return [pop[int(i)] for i in idx]


def _get_random_choice_vanilla(num_items, max_items=4):
    """utils.py:21-27: uniform over 1..min(num_items, max_items)."""
    return pick_index(min(num_items, max_items)) + 1


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

import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    """PFANN_TEST_ORDER=reverse | shuffle:<seed> runs the tests in another order: the library keeps per-process
    caches (kernel attributes, workspaces), so order dependence is a real failure mode worth checking."""
    order = os.environ.get('PFANN_TEST_ORDER', '')
    if order == 'reverse':
        items.reverse()
    elif order.startswith('shuffle'):
        import random
        random.Random(int(order.split(':')[1]) if ':' in order else 0).shuffle(items)

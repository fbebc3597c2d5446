"""Import stub (legged_gym/utils/logger.py:4-5 imports matplotlib for offline plots). Oracle tooling only."""


def use(*a, **k):
    pass

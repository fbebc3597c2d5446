"""Import stub (legged_gym/utils/logger.py:6). Oracle tooling only."""

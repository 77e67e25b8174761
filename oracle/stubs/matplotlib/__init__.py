"""Empty stand-in so that the reference's unconditional `import matplotlib.pyplot`
(toppra/algorithm/algorithm.py:14) succeeds on boxes without matplotlib.
TEST INFRASTRUCTURE ONLY (used by oracle/ref_loader.py)."""

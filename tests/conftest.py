import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the authoring container exposes 8 vCPUs but schedules about one; keep the C oracle's OpenMP team small
os.environ.setdefault('OMP_NUM_THREADS', '2')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')

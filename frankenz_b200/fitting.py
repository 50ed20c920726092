"""Re-export of the estimators (mirrors frankenz/fitting.py:23-29).

SelfOrganizingMap / GrowingNeuralGas are out of scope of this build.
"""
from .bruteforce import BruteForce
from .knn import NearestNeighbors

__all__ = ["BruteForce", "NearestNeighbors"]

"""Re-export of the estimators (mirrors frankenz/fitting.py:23-29).

The training of SelfOrganizingMap / GrowingNeuralGas is out of scope of this build (sequential
online updates); the stages of those classes that score photometry against the trained nodes run on the GPU through
`frankenz_b200.networks.NetworkFit`.
"""
from .bruteforce import BruteForce
from .knn import NearestNeighbors
from .networks import NetworkFit

__all__ = ["BruteForce", "NearestNeighbors", "NetworkFit"]

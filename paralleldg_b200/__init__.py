"""paralleldg_b200: B200-native implementation of parallelDG's parallel Metropolis-Hastings hot
path (junction-tree proposals, clique scorers, accept/reject) for many chains at once."""
__version__ = "0.1.0"

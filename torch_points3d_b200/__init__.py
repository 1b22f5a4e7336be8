"""torch_points3d_b200 -- B200-native (sm_100a) sampling / grouping / neighbour-aggregation
operators behind the torch-points3d operator API.  See DESIGN.md.

Importing the package does not touch CUDA and does not load the shared library (DataLoader
workers may import it); the first operator call does, and fails loudly if the library is not
built or no CUDA device is present."""
__version__ = "0.1.0"

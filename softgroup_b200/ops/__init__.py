from .functions import *  # noqa: F401,F403
from .functions import bfs_cluster_segments  # noqa: F401

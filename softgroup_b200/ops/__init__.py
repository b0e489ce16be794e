from .functions import *  # noqa: F401,F403
from .functions import ballquery_batch_p_nosync, bfs_cluster_segments, build_octree, gather_rows, group_entries  # noqa: F401

"""Partition plumbing and on-disk formats either side of the rasterizer hot path: COLMAP model
I/O, the VastGaussian scene partition of split_scene.py (CPU only), box.txt, Gaussian PLY files, and a CPU distCUDA2."""
from .colmap_io import (Camera, Image, Point3D, qvec2rotmat, read_model, rotmat2qvec, write_model)  # noqa: F401
from .knn_cpu import dist2_knn3_cpu  # noqa: F401
from .vast import (camera_position_based_region_division, coverage_based_point_selection, list_tiles, partition_scene,  # noqa: F401
                   position_based_data_selection, read_box, split_scene, tiles_for_rank, transform_colmap,
                   visibility_based_camera_selection, write_box, write_tiles)

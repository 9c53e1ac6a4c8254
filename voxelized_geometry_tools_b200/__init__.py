"""B200-native (sm_100a) backend for the occupancy -> SDF hot path of voxelized_geometry_tools.

Public host-side surface (mirrors the reference's names for this path):

    OccupancyMap, SignedDistanceField, SignedDistanceFieldGenerationParameters, VoxelGridSizes
    PointCloudVoxelizationFilterOptions, VectorPointCloudWrapper, B200PointCloudVoxelizer, ...

``voxelized_geometry_tools_b200.device`` (imports torch) holds the device-resident entry points and
``voxelized_geometry_tools_b200.sharded`` the multi-GPU slab-sharded SDF. Everything computes on
the GPU through libvgt_b200.so (include/vgt_b200.h); there is no CPU fallback.
"""
from ._capi import BackendUnavailable, device_count  # noqa: F401
from .grids import (  # noqa: F401
    ComputeDistanceFieldTransformInPlace,
    ComputeSquaredDistanceFields,
    ExtractSignedDistanceFieldFromMask,
    OccupancyComponentMap,
    OccupancyMap,
    SignedDistanceField,
    SignedDistanceFieldGenerationParameters,
    TaggedObjectOccupancyComponentMap,
    TaggedObjectOccupancyMap,
    VoxelGridSizes,
)
from .pointcloud_voxelization import (  # noqa: F401
    B200PointCloudVoxelizer,
    Float32PointCloudWrapper,
    GetAvailableBackends,
    MakePointCloudVoxelizer,
    PointCloudVoxelizationFilterOptions,
    PointCloudVoxelizationInterface,
    PointCloudWrapper,
    SeenAs,
    VectorPointCloudWrapper,
    VoxelizerRuntime,
)

__version__ = "0.1.0"

import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import voxelized_geometry_tools_b200 as vgt
from oracle import reference_oracle
rng = np.random.default_rng(1)
for shape in [(9,1,1),(1,9,1),(1,1,9),(9,1,7),(9,7,1),(1,9,7),(9,5,7)]:
    for scale, inf_rate in ((1,0.0),(400,0.3),(7,1.0),(3,0.9)):
        field = rng.integers(0, scale+1, size=shape).astype(np.float64)
        field[rng.random(shape) < inf_rate] = np.inf
        want = reference_oracle.transform_inplace(field.copy())
        got = vgt.ComputeDistanceFieldTransformInPlace(field.copy())
        ok = np.array_equal(got, want)
        print(shape, scale, inf_rate, 'OK' if ok else 'FAIL')
        if not ok:
            print(' in  ', field.reshape(-1)[:20]); print(' got ', got.reshape(-1)[:20]); print(' want', want.reshape(-1)[:20])

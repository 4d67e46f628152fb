/* CPU oracle: restatement of spconv's serial points->voxels generator.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED: spconv (docs pin v1.2.1, /root/reference/docs/md_files/installation.md:45-47)
 * is a third-party dependency that is neither vendored under /root/reference nor installed in
 * this image, and the reference has no test or golden vector at this boundary.  This file restates
 * the published `points_to_voxel_3d_np` algorithm exactly as the reference consumes it:
 *   /root/reference/opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:41-43  grid = round((max-min)/vs)
 *   ... :46-60   VoxelGenerator(voxel_size, point_cloud_range, max_num_points, max_voxels)
 *   ... :62-85   preprocess(): voxels (M,32,4) f32, coordinates (M,3) i32 [z,y,x], num_points (M,) i32
 *   ... :145-174 collate: prepend the agent index -> coords (M,4) [a,z,y,x]
 *
 * Semantics: points are visited in input order; c_j = floor((p_j - min_j)/vs_j) in
 * float32; a point is dropped if any c_j is outside [0,grid_j); voxels are numbered in order of
 * first appearance; once max_voxels voxels exist, points that would open a new voxel are skipped
 * (points falling into existing voxels are still appended); a voxel keeps its first max_pts points.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* table: caller-provided scratch of grid[0]*grid[1]*grid[2] int32, any content (reset here).
 * voxels must be zero-filled by the caller?  No: we zero-fill rows as voxels are opened.
 * returns the number of voxels. */
int oracle_voxelize(const float *pts, int n_pts, int n_feat,
                    const float range[6], const float vsize[3], const int grid[3],
                    int max_pts, int max_voxels,
                    float *voxels, int32_t *coords_zyx, int32_t *num_points, int32_t *table)
{
    const long cells = (long)grid[0] * grid[1] * grid[2];
    for (long i = 0; i < cells; ++i) table[i] = -1;
    int n_vox = 0;
    for (int i = 0; i < n_pts; ++i) {
        const float *p = pts + (long)i * n_feat;
        int c[3];
        int ok = 1;
        for (int j = 0; j < 3; ++j) {
            /* float32 arithmetic, as the templated DType=float generator does */
            float f = floorf((p[j] - range[j]) / vsize[j]);
            if (!(f >= 0.0f) || !(f < (float)grid[j])) { ok = 0; break; }
            c[j] = (int)f;
        }
        if (!ok) continue;
        long cell = ((long)c[2] * grid[1] + c[1]) * grid[0] + c[0];   /* z,y,x major->minor */
        int v = table[cell];
        if (v < 0) {
            if (n_vox >= max_voxels) continue;
            v = n_vox++;
            table[cell] = v;
            coords_zyx[3 * v + 0] = c[2];
            coords_zyx[3 * v + 1] = c[1];
            coords_zyx[3 * v + 2] = c[0];
            num_points[v] = 0;
            memset(voxels + (long)v * max_pts * n_feat, 0, sizeof(float) * max_pts * n_feat);
        }
        if (num_points[v] < max_pts) {
            memcpy(voxels + ((long)v * max_pts + num_points[v]) * n_feat, p, sizeof(float) * n_feat);
            num_points[v] += 1;
        }
    }
    return n_vox;
}

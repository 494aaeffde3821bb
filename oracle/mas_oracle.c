/* CPU oracle for Monotonic Alignment Search -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's Cython kernel
 *   /root/reference/model/monotonic_align/core.pyx:9-35  (maximum_path_each)
 *   /root/reference/model/monotonic_align/core.pyx:40-45 (maximum_path_c)
 * Pinned against (a) the reference itself, cythonized from where it lies into oracle/_ref/
 * by oracle/build_oracle.py, and (b) the committed fixtures tests/golden/mas_*.npz, which
 * were produced by the reference's own maximum_path (tests/golden/make_golden.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may call this.
 *
 * All arithmetic is IEEE float32 add/compare exactly as the generated C does it
 * (max() lowers to (v_prev > v_cur) ? v_prev : v_cur); build with -O2 -ffp-contract=off.
 */
#include <stdint.h>

static void mas_each(int32_t *path, float *value, int t_x, int t_y, int ld, float max_neg_val)
{
    int x, y;
    int index = t_x - 1;
    for (y = 0; y < t_y; y++) {                                   /* core.pyx:17-30 */
        int lo = t_x + y - t_y; if (lo < 0) lo = 0;
        int hi = (t_x < y + 1) ? t_x : (y + 1);
        for (x = lo; x < hi; x++) {
            float v_cur, v_prev;
            if (x == y) v_cur = max_neg_val; else v_cur = value[x * ld + y - 1];
            if (x == 0) v_prev = (y == 0) ? 0.0f : max_neg_val;
            else        v_prev = value[(x - 1) * ld + y - 1];
            value[x * ld + y] = ((v_prev > v_cur) ? v_prev : v_cur) + value[x * ld + y];
        }
    }
    for (y = t_y - 1; y >= 0; y--) {                              /* core.pyx:32-35 */
        path[index * ld + y] = 1;
        if (index != 0 && (index == y || value[index * ld + y - 1] < value[(index - 1) * ld + y - 1]))
            index = index - 1;
    }
}

/* paths, values: [b][t_x_max][t_y_max] C-contiguous; values is modified in place (as in the
 * reference, which works on its private numpy copy). */
void mas_oracle_maximum_path_c(int32_t *paths, float *values, const int32_t *t_xs, const int32_t *t_ys,
                               int b, int t_x_max, int t_y_max, float max_neg_val)
{
    int i;
    for (i = 0; i < b; i++)                                       /* core.pyx:44 (serial prange) */
        mas_each(paths + (int64_t)i * t_x_max * t_y_max, values + (int64_t)i * t_x_max * t_y_max,
                 t_xs[i], t_ys[i], t_y_max, max_neg_val);
}

/* CPU oracle kernels in plain C (TEST INFRASTRUCTURE -- see oracle/__init__.py).
 *
 * oracle_knn_exact: the definition of "exact kNN" the CUDA path must match bit for bit
 * (SURVEY.md 8(c)-iii).  It replaces hnswlib 0.8.0 `Index(space="l2").knn_query`
 * (scarf/ann.py:14-28,201-205: float32 vectors, SQUARED L2, ascending) followed by
 * `fix_knn_query` (scarf/ann.py:31-52: drop the self hit):
 *     d(a,b) = (float) sum_{t=0..dim-1} ((double)a[t] - (double)b[t])^2      (t ascending, no FMA)
 *     neighbours ordered by (d, index), the query's own row excluded when self_offset >= 0.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC  (no -ffast-math: the sum order is the spec).
 */
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int better(float d, int64_t i, float dk, int64_t ik) { return d < dk || (d == dk && i < ik); }

int32_t oracle_knn_exact(const float* q, int64_t nq, const float* ref, int64_t nref, int32_t dim, int32_t k,
                         int64_t self_offset, int64_t* out_idx, float* out_dist, int32_t nthreads) {
  if (k <= 0 || dim <= 0 || nq < 0 || nref <= 0) return 1;
  if ((self_offset >= 0 ? nref - 1 : nref) < k) return 2;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < nq; ++i) {
    const float* a = q + i * (int64_t)dim;
    int64_t* bi = out_idx + i * (int64_t)k;
    float* bd = out_dist + i * (int64_t)k;
    int32_t cnt = 0;
    const int64_t self = self_offset >= 0 ? i + self_offset : -1;
    for (int64_t j = 0; j < nref; ++j) {
      if (j == self) continue;
      const float* b = ref + j * (int64_t)dim;
      double acc = 0.0;
      for (int32_t t = 0; t < dim; ++t) {
        double df = (double)a[t] - (double)b[t];
        acc = acc + df * df;
      }
      float d = (float)acc;
      if (cnt == k && !better(d, j, bd[k - 1], bi[k - 1])) continue;
      int32_t p = cnt < k ? cnt : k - 1;
      while (p > 0 && better(d, j, bd[p - 1], bi[p - 1])) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
      bd[p] = d; bi[p] = j;
      if (cnt < k) ++cnt;
    }
  }
  return 0;
}

/*
 * pm_oracle.c -- the reference's photon-map radiance estimate as a callable (TEST INFRASTRUCTURE).
 *
 * Linked with the UNMODIFIED /root/reference/src/libs/photon_map/pm.c into oracle/_ref/libpm_ref.so by
 * oracle/build_ref.py.  Only tests/ load it: tests/test_gpu_knn.py hands the same photons to this library and to the
 * device map (frt_photons_import / frt_photons_finish) and compares pm_irradiance_estimate (pm.c:91-156, through the
 * reference's own left-balanced kd-tree, pm_balance :329 and pm_locate_photons :163) with frt_photons_estimate.
 */
#include <stdlib.h>
#include <string.h>

#include "src/libs/photon_map/pm.h"

/* photons arrive in the device's export layout split into arrays: pos / power n x 3 floats, theta / phi bytes */
int
pm_oracle_estimate(long n_photons, const float *pos, const float *power, const unsigned char *theta, const unsigned char *phi,
                   long n_queries, const double *qpos, const double *qnormal, double radius, int nphotons, double cone_k,
                   double *irrad, long *found, float *lost_pos, int *n_lost)
{
    PhotonMap pm;
    init_Photon_map(n_photons, &pm);
    for (long i = 0; i < n_photons; ++i) {
        Photon *p = &pm.photons[i + 1]; /* pm_store (pm.c:261-301) fills 1..stored */
        memset(p, 0, sizeof(*p));
        for (int k = 0; k < 3; ++k) {
            p->pos[k] = (double)pos[3 * i + k];
            p->power[k] = (double)power[3 * i + k];
            if (p->pos[k] < pm.bbox_min[k]) pm.bbox_min[k] = p->pos[k];
            if (p->pos[k] > pm.bbox_max[k]) pm.bbox_max[k] = p->pos[k];
        }
        p->theta = theta[i];
        p->phi = phi[i];
    }
    pm.stored_photons = n_photons;
    pm_balance(&pm);
    /* pm_locate_photons only descends from nodes with index < half_stored_photons = stored / 2 - 1 (pm.c:173, :372), so the
     * children of the last one or two inner nodes -- heap slots 2 * half_stored_photons .. stored, three or four photons --
     * are never looked at by any query.  Report them (up to 4 positions) so that a caller can tell which queries they touch. */
    if (n_lost != NULL) {
        int k = 0;
        for (long i = 2 * pm.half_stored_photons; i <= pm.stored_photons && lost_pos != NULL; ++i) {
            if (i >= 1 && i / 2 >= pm.half_stored_photons && k < 4) {
                lost_pos[3 * k] = (float)pm.photons[i].pos[0];
                lost_pos[3 * k + 1] = (float)pm.photons[i].pos[1];
                lost_pos[3 * k + 2] = (float)pm.photons[i].pos[2];
                ++k;
            }
        }
        *n_lost = k;
    }
    for (long q = 0; q < n_queries; ++q) {
        double p[3] = { qpos[3 * q], qpos[3 * q + 1], qpos[3 * q + 2] };
        double n[3] = { qnormal[3 * q], qnormal[3 * q + 1], qnormal[3 * q + 2] };
        found[q] = pm_irradiance_estimate(&pm, irrad + 3 * q, p, n, radius, nphotons, cone_k);
    }
    delete_Photon_map(&pm);
    return 0;
}

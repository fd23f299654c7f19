/* host_driver.c -- a compiled (plain C) host driving the C ABI in exactly the order the Fortran shim
 * fortran/mod_particleMover_b200.f90 does (ipic_attach -> ipic_push_state -> depositJ -> moveParticles).
 * Stands in for the Fortran host, which cannot be compiled in this image (no Fortran compiler).
 *
 * usage: host_driver <in.bin> <out.bin> [level3]
 * in.bin : int32 scheme, Nx, smooth, nsp, T, finalIdx ; double del_t, Vl, Vr, rf ;
 *          double grid[Nx], dx[ndx] ; int32 bc[nbc] ; double phi[Nx], phi_f[Nx] ;
 *          per species: double mass, q, w_p ; int32 N_p[T] ; double phaseSpace[T][finalIdx][4]
 * out.bin: double J[Nh] ; per species: int32 N_p[T] ; double phaseSpace[T][finalIdx][4] ; ipic_move_stats
 * With a third argument (level-3 integration) what mod_simulation.f90:161-168 and a diagnostics dump run
 * between two steps happens on the device before the download: addMaxwellianLostParticles (T_e = 3, T_i = 0.1 eV), loadParticleDensity and
 * the EEDF extrema of the electrons; out.bin then ends with int64 numAdded ; double energyAdd ; double density[Nx] ; double extrema[3].
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/ipic1d_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != IPIC_OK) {                                                            \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ipic_last_error());            \
            return 10 + rc_;                                                             \
        }                                                                                \
    } while (0)

static int rd(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n ? 0 : 1; }

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    int32_t hdr[6];
    double sc[4];
    if (rd(fi, hdr, sizeof hdr) || rd(fi, sc, sizeof sc)) return 4;
    const int scheme = hdr[0], Nx = hdr[1], smooth = hdr[2], nsp = hdr[3], T = hdr[4], finalIdx = hdr[5];
    const int Nh = scheme == IPIC_SCHEME_NGP ? Nx - 1 : Nx + 1;
    const int ndx = scheme == IPIC_SCHEME_NGP ? Nh : Nx, nbc = scheme == IPIC_SCHEME_NGP ? Nx : Nh;
    double *grid = malloc(sizeof(double) * Nx), *dx = malloc(sizeof(double) * ndx), *phi = malloc(sizeof(double) * Nx),
           *phi_f = malloc(sizeof(double) * Nx), *J = malloc(sizeof(double) * Nh);
    int32_t *bc = malloc(sizeof(int32_t) * nbc);
    if (rd(fi, grid, sizeof(double) * Nx) || rd(fi, dx, sizeof(double) * ndx) || rd(fi, bc, sizeof(int32_t) * nbc) ||
        rd(fi, phi, sizeof(double) * Nx) || rd(fi, phi_f, sizeof(double) * Nx))
        return 5;

    ipic_handle h;
    CHECK(ipic_create(&h, scheme, Nx, grid, dx, bc, smooth, 0));
    CHECK(ipic_init_solver(h, sc[1], sc[2], sc[3], 0.0));
    const size_t psz = (size_t)T * finalIdx * 4;
    double **ps = malloc(sizeof(double *) * nsp);
    int32_t **Np = malloc(sizeof(int32_t *) * nsp);
    for (int s = 0; s < nsp; ++s) {
        double msq[3];
        int32_t idx;
        if (rd(fi, msq, sizeof msq)) return 6;
        CHECK(ipic_add_species(h, msq[0], msq[1], msq[2], &idx));
        Np[s] = malloc(sizeof(int32_t) * T);
        CHECK(ipic_host_alloc((void **)&ps[s], psz * sizeof(double)));   /* pinned phaseSpace, as INTEGRATION.md recommends */
        if (rd(fi, Np[s], sizeof(int32_t) * T) || rd(fi, ps[s], psz * sizeof(double))) return 7;
    }
    fclose(fi);

    /* ipic_push_state */
    CHECK(ipic_set_field(h, IPIC_F_PHI, phi, Nx));
    CHECK(ipic_set_field(h, IPIC_F_PHI_F, phi_f, Nx));
    for (int s = 0; s < nsp; ++s) CHECK(ipic_upload_particles(h, s, ps[s], Np[s], T, finalIdx));
    /* depositJ */
    CHECK(ipic_deposit_J(h, sc[0]));
    CHECK(ipic_get_field(h, IPIC_F_J, J, Nh));
    /* moveParticles */
    CHECK(ipic_move_particles(h, sc[0]));
    int64_t numAdded = 0;
    double energyAdd = 0.0, extrema[3] = {0, 0, 0};
    double *dens = calloc((size_t)Nx, sizeof(double));
    if (argc > 3) {
        CHECK(ipic_set_injection_seed(h, 2026u));
        CHECK(ipic_add_maxwellian_lost_particles(h, 3.0, 0.1, &energyAdd, &numAdded));
        CHECK(ipic_load_particle_density(h, 0, dens));
        CHECK(ipic_velocity_extrema(h, 0, extrema));
    }

    FILE *fo = fopen(argv[2], "wb");
    if (!fo) return 8;
    fwrite(J, sizeof(double), Nh, fo);
    for (int s = 0; s < nsp; ++s) {
        ipic_move_stats st;
        memset(ps[s], 0, psz * sizeof(double));
        CHECK(ipic_download_particles(h, s, ps[s], Np[s], T, finalIdx));
        CHECK(ipic_get_move_stats(h, s, &st));
        fwrite(Np[s], sizeof(int32_t), T, fo);
        fwrite(ps[s], sizeof(double), psz, fo);
        fwrite(&st, sizeof st, 1, fo);
        CHECK(ipic_host_free(ps[s]));
    }
    if (argc > 3) {
        fwrite(&numAdded, sizeof numAdded, 1, fo);
        fwrite(&energyAdd, sizeof energyAdd, 1, fo);
        fwrite(dens, sizeof(double), (size_t)Nx, fo);
        fwrite(extrema, sizeof(double), 3, fo);
    }
    fclose(fo);
    CHECK(ipic_destroy(h));
    printf("host_driver ok\n");
    return 0;
}

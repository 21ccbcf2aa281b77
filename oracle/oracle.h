/*
 * oracle.h — C interface of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libdxmc_b200.so) never does.
 *
 * PARITY UNPINNED: the arithmetic of the hot path lives in medicalphysics/DXMClib (GIT_TAG
 * develop, a floating branch, R:CMakeLists.txt:71-76), which is not in /root/reference and cannot
 * be built offline; the reference's own tests pin nothing at this boundary.
 * The oracle therefore follows (i) the parts verified from OpenDXMC's sources and (ii) the
 * recalled DXMClib design intent, and is pinned only against analytic known answers
 * (tests/test_oracle_*.py) and the Random123 Philox known-answer vectors.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "../include/dxb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world orc_world;

typedef struct orc_stats {
    uint64_t histories, steps, interactions, deposits;
    double energy_emitted_kev, energy_deposited_kev;
    double calibration_factor;
    double seconds;
    int threads;
    uint64_t hops; /* slab-local majorants: tentative steps that ended on a slab face */
} orc_stats;

/* Device mirroring (default on): the oracle quantises its inputs where the device stores a quantised copy (24-bit voxel
 * density, f32 majorant, bowtie knots, alias acceptance values, shell constants, exposure geometry).  0 switches all
 * of it off - pure f64 on the caller's data; process-wide, set it BEFORE orc_world_create. */
void orc_set_device_mirroring(int on);
int orc_get_device_mirroring(void);

/* World<AAVoxelGrid>: copies everything it needs (tables are deep-copied). */
orc_world* orc_world_create(const uint64_t dim[3], const double spacing_cm[3], const double* density,
                            const uint8_t* material, uint32_t n_materials, const dxb_material_tables* tables);
void orc_world_destroy(orc_world*);
/* Slab-local majorants (the kernel's LM builds): track with the table the device built (dxb_get_local_majorant; n_slabs * 16
 * floats; n_slabs < 2 switches it off) or with the oracle's own f64 table for slabs of 2^shift voxel layers (returns the
 * number of slabs). */
void orc_world_set_local_majorant(orc_world*, int shift, int n_slabs, const float* ratio);
int orc_world_build_local_majorant(orc_world*, int shift);
/* Dense-box tracking (the kernel's DB builds): faces = {x0, y0, z0, x1, y1, z1} [cm] and the 16 outside ratios the device built
 * (dxb_get_dense_box), NULL switches it off; or the oracle's own box from its f64 tables (thin voxel: attenuation <= theta x
 * majorant at every energy; returns 1 if the box is a proper part of the grid). */
void orc_world_set_dense_box(orc_world*, const float* faces, const float* ratio);
int orc_world_build_dense_box(orc_world*, double theta);
/* air + PMMA tables for the nested CTDI calibration and the DAP / air-kerma calibrations */
void orc_world_set_reference_materials(orc_world*, const dxb_material_tables* air, const dxb_material_tables* pmma,
                                       double air_density, double pmma_density);

/* Transport::operator(): clears the energy tallies, runs every history of the beam on n_threads
 * std::threads (0 = hardware_concurrency), optionally restricted to shard (rank, world) with the same
 * 65536-history round-robin blocks as the product; energy/energy_sq/n_events (may be NULL) receive the
 * per-voxel tallies of this beam. */
int orc_run(orc_world*, const dxb_beam_desc*, int physics_mode, uint64_t seed, int n_threads, uint64_t rank, uint64_t world,
            double* energy, double* energy_sq, uint64_t* n_events, orc_stats* stats);
/* full transport(): tallies + calibration factor + dose accumulation into dose/variance/events (caller zeroes) */
int orc_transport(orc_world*, const dxb_beam_desc*, int physics_mode, int use_beam_calibration, uint64_t seed,
                  uint64_t calibration_histories, int n_threads, double* dose, double* variance, uint64_t* n_events,
                  orc_stats* stats);
double orc_ct_calibration(orc_world*, const dxb_beam_desc*, int physics_mode, uint64_t seed, uint64_t calibration_histories,
                          int n_threads);

/* building blocks exposed for unit tests */
void orc_philox4x32_10(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]);
uint64_t orc_beam_number_of_exposures(const dxb_beam_desc*);
int orc_beam_exposure(const dxb_beam_desc*, uint64_t index, dxb_exposure* out);
double orc_bowtie_weight(const dxb_bowtie*, double angle);
double orc_aec_weight(const dxb_aec*, const double pos[3]);
double orc_organ_aec_weight(const dxb_organ_aec*, double angle);
void orc_attenuation(const dxb_material_tables*, double energy_kev, double out_pict[4]); /* photo, incoh, coh, total */
double orc_majorant(const orc_world*, double energy_kev);
/* sample n Compton / Rayleigh events at fixed energy; writes cos(theta) and E'/E */
void orc_sample_compton(const dxb_material_tables*, int physics_mode, double energy_kev, uint64_t seed, uint64_t n,
                        double* cos_theta, double* energy_ratio);
void orc_sample_rayleigh(const dxb_material_tables*, int physics_mode, double energy_kev, uint64_t seed, uint64_t n,
                         double* cos_theta);
void orc_sample_source(const dxb_beam_desc*, uint64_t seed, uint64_t first_history, uint64_t n,
                       double* pos3, double* dir3, double* energy, double* weight);
/* per-organ dose, R:src/libopendxmc/dosetablepipeline.cpp:60-84 */
void orc_organ_dose(const double* dose, const double* density, const uint8_t* organ, uint64_t n, double voxel_volume,
                    uint32_t n_organs, double* dose_out, double* mass_out, uint64_t* count_out);
/* reference post-processing R:src/libopendxmc/simulationpipeline.cpp:174-232; returns 1 if units are uGy */
int orc_postprocess(double* dose, double* variance, double* events, const uint8_t* material, uint64_t n, int delete_air);
/* CT segmentation R:src/libopendxmc/ctsegmentationpipeline.cpp:131-156 given thresholds and attenuations */
void orc_segment(const double* hu, uint64_t n, const double* sep, int n_sep, const double* mat_att, double water_att_dens,
                 double air_att_dens, uint8_t* material, double* density);

#ifdef __cplusplus
}
#endif
#endif

/*
 * dxb.h — C ABI of libdxmc_b200: the B200-native replacement for the DXMClib
 * `dxmc::Transport` photon-history path that OpenDXMC drives.
 *
 * Citation convention: R:path:line = /root/reference/path:line (OpenDXMC).
 *
 * OpenDXMC has no FFI layer for this path: it instantiates DXMClib's C++
 * templates inside libopendxmc (R:src/libopendxmc/simulationpipeline.cpp:124-235).
 * This header is the boundary a replacement must export; the C++ shim headers
 * under include/dxmc/ re-create the consumed C++ names on top of it, and
 * opendxmc_b200/ (Python, ctypes) mirrors the same names for tests and bench.
 *
 * Rules of the ABI
 *  - plain C, no exceptions cross it; every call returns a dxb_status (or a value).
 *  - caller owns every array it passes in (the reference hands `const&` to
 *    DataContainer vectors, R:src/libopendxmc/simulationpipeline.cpp:147-149);
 *    outputs are caller-allocated.
 *  - one dxb_run at a time per context; progress/stop are lock-free and may be
 *    touched from other threads (R:src/libopendxmc/simulationpipeline.cpp:112-122,259-262).
 *  - there is NO CPU fallback: without a CUDA device dxb_create fails with DXB_ECUDA.
 *
 * Units follow the reference: lengths cm (R:src/libopendxmc/datacontainer.cpp:241-253),
 * energies keV, density g/cm3, angles rad, dose mGy.
 */
#ifndef DXB_H
#define DXB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DXB_ABI_VERSION 2

typedef enum dxb_status {
    DXB_OK = 0,
    DXB_EINVAL = 1,    /* bad argument */
    DXB_EMATERIAL = 2, /* Material::byWeight -> nullopt (R:src/libopendxmc/simulationpipeline.cpp:136-141) */
    DXB_ECUDA = 3,     /* CUDA error / no device / extension not usable */
    DXB_ESTATE = 4,    /* call order violated (e.g. run before set_grid) */
    DXB_ECANCELLED = 5,/* stop flag observed; the reference conflates this with "finished" (:119-121,:234) */
    DXB_ENOMEM = 6
} dxb_status;

/* ---- physics modes: the CORRECTION template argument of AAVoxelGrid<5,CORRECTION,255>
 *      (R:src/libopendxmc/simulationpipeline.cpp:127,247-255; GUI default 1,
 *      R:src/libopendxmc/simulationwidget.cpp:70-75) */
#define DXB_PHYSICS_NONE 0      /* free-electron Klein-Nishina, Thomson Rayleigh */
#define DXB_PHYSICS_LIVERMORE 1 /* scatter function S(q) and form factor F(q) */
#define DXB_PHYSICS_IA 2        /* impulse approximation: shell binding, Doppler, fluorescence */

/* ====================================================================== */
/* Materials  (dxmc::Material<5>, dxmc::NISTMaterials, dxmc::AtomHandler)  */
/* ====================================================================== */

typedef struct dxb_material dxb_material; /* opaque, immutable after creation */

/* Material<5>::byWeight(map<Z,weight>) — R:src/libopendxmc/simulationpipeline.cpp:136.
 * Weights need not be normalised (ICRP media are mass-%, the appended air is a
 * fraction, R:src/libopendxmc/icrpphantomimportpipeline.cpp:273,300).
 * Fails with DXB_EMATERIAL for Z outside 1..92, non-positive total weight, n == 0. */
int dxb_material_by_weight(dxb_material** out, uint32_t n, const uint32_t* Z, const double* weight);
/* Material<5>::byNistName — R:src/libopendxmc/ctsegmentationpipeline.cpp:73-75,83-85 */
int dxb_material_by_nist_name(dxb_material** out, const char* name);
/* Material<5>::byChemicalFormula ("H2O", "C5O2H8") */
int dxb_material_by_chemical_formula(dxb_material** out, const char* formula);
void dxb_material_destroy(dxb_material*);

/* attenuationValues(E): mass attenuation [cm2/g] {photoelectric, incoherent, coherent}.
 * `physics_mode` 0 gives the free-electron incoherent and Thomson-type coherent
 * values that mode samples from; 1 and 2 the form-factor/scatter-function ones. */
int dxb_material_attenuation(const dxb_material*, double energy_kev, double out_pic[3]);
double dxb_material_mass_energy_transfer(const dxb_material*, double energy_kev); /* mu_tr/rho [cm2/g] */
double dxb_material_effective_z(const dxb_material*);
double dxb_material_form_factor(const dxb_material*, double x_inv_angstrom);     /* F(x), x = sin(theta/2)/lambda */
double dxb_material_scatter_factor(const dxb_material*, double x_inv_angstrom);  /* S(x)/Z_total in [0,1] */

/* dxmc::NISTMaterials — R:src/libopendxmc/ctsegmentationpipeline.cpp:74-76,
 * R:src/libopendxmc/otherphantomimportpipeline.cpp:87-93 */
int dxb_nist_count(void);
const char* dxb_nist_name(int index);
double dxb_nist_density(const char* name); /* g/cm3, <0 if unknown */
/* composition: returns element count, fills up to cap entries (mass fractions) */
int dxb_nist_composition(const char* name, uint32_t* Z, double* weight, int cap);
/* dxmc::AtomHandler::toSymbol(Z) — R:src/libopendxmc/hdf5wrapper.cpp:429 */
const char* dxb_atom_symbol(uint32_t Z);
double dxb_atom_weight(uint32_t Z);
double dxb_atom_standard_density(uint32_t Z);

/* The interaction tables of one material, exactly as the device consumes them
 * (double precision master copy).  This is also what tests hand to the CPU oracle
 * so that both sides run on identical data.
 * All arrays are owned by the material. */
#define DXB_MAX_SHELLS 5
typedef struct dxb_shell {
    double binding_energy_kev;
    double n_electrons;         /* electrons per "molecule-average atom" in this shell group */
    double n_electrons_fraction;/* share of all electrons of the material */
    double photo_fraction_above;/* probability that a photoelectric event above the edge ionises this shell */
    double fluor_yield;         /* probability of a fluorescence photon after ionisation */
    double fluor_energy_kev;    /* mean line energy */
    double compton_j0;          /* Compton profile J(0) [1/(m_e c alpha)] for the analytic profile */
} dxb_shell;

typedef struct dxb_material_tables {
    uint32_t n_energy;      /* nodes of the semi-log energy grid: node(i) = e_min * 2^(i/P) * (1 + (i%P)/P),
                               P = nodes_per_octave_e (uniform spacing inside each octave) */
    double   e_min_kev, e_max_kev;
    const double* photo;    /* [n_energy] cm2/g */
    const double* incoh;    /* [n_energy] Livermore (S(q)-weighted) */
    const double* coh;      /* [n_energy] form-factor Rayleigh */
    const double* incoh_kn; /* [n_energy] free-electron Klein-Nishina (mode 0) */
    const double* coh_thomson; /* [n_energy] mode-0 value (== coh: DXMClib keeps the cross-section, changes only the angular law) */
    const double* etr;      /* [n_energy] mass energy-transfer coefficient (kerma; CT/DX calibration) */
    uint32_t n_x;           /* nodes of the semi-log momentum-transfer grid x [1/Angstrom], P = nodes_per_octave_x */
    double   x_min, x_max;
    const double* ff_cdf;   /* [n_x] A(x_k) = integral_0^{x_k^2} F(t)^2 d(t) / Z^2, t = x^2 */
    const double* sf;       /* [n_x] S(x_k)/Z */
    uint32_t n_shells;      /* <= DXB_MAX_SHELLS, most tightly bound first */
    dxb_shell shells[DXB_MAX_SHELLS];
    double   rest_electrons_fraction; /* electrons not covered by `shells`: one unbound group (mode 2) */
    double   rest_compton_j0;         /* J(0) of that group (electron-weighted mean of its orbitals); 0: electrons at rest */
    double   electrons_per_gram;
    double   effective_z;
    uint32_t nodes_per_octave_e, nodes_per_octave_x;
} dxb_material_tables;
int dxb_material_tables_get(const dxb_material*, dxb_material_tables* out);
/* The reverse direction - the drop-in route for externally built physics data (DXMClib ships EPICS2014-derived tables,
 * R:src/app/CMakeLists.txt:38 `dxmclib_add_physics_list`): a material made of caller-supplied tables on the library's
 * grids (geometry as returned by dxb_material_tables_get; DXB_EINVAL otherwise, DXB_EMATERIAL for negative / non-finite
 * values or a non-monotone ff_cdf).  The arrays are copied; host lookups, the device tables and the majorant follow
 * them, nothing is recomputed from the built-in atom model.  incoh_kn may be NULL (then = incoh). */
int dxb_material_from_tables(dxb_material** out, const dxb_material_tables* tables);
/* library-wide table geometry */
uint32_t dxb_table_n_energy(void);
double   dxb_table_e_min(void);
double   dxb_table_e_max(void);

/* ====================================================================== */
/* Tube  (dxmc::Tube) — R:src/libopendxmc/ctsegmentationpipeline.cpp:66-71, */
/*                       R:src/libopendxmc/beamsettingsmodel.cpp:257-346    */
/* ====================================================================== */
#define DXB_TUBE_MAX_FILT 8
typedef struct dxb_tube_desc {
    double voltage_kv;       /* clamped to [20,150] like the GUI */
    double anode_angle_deg;  /* default 12 */
    double energy_resolution_kev; /* default 1 */
    uint32_t n_filt;
    uint32_t filt_Z[DXB_TUBE_MAX_FILT];
    double   filt_mm[DXB_TUBE_MAX_FILT];
} dxb_tube_desc;
/* Tube::getEnergy(): number of bins; fills up to cap */
int dxb_tube_energies(const dxb_tube_desc*, double* energy_kev, int cap);
/* Tube::getSpecter(energies, normalize) */
int dxb_tube_spectrum(const dxb_tube_desc*, const double* energy_kev, int n, int normalize, double* weight);
double dxb_tube_mean_energy(const dxb_tube_desc*);
double dxb_tube_al_half_value_layer_mm(const dxb_tube_desc*);

/* ====================================================================== */
/* Beams                                                                    */
/* ====================================================================== */
typedef enum dxb_beam_type {
    DXB_BEAM_DX = 0,            /* dxmc::DXBeam<false> (+ OpenDXMC subclass, R:src/libopendxmc/dxmc_specialization.cpp:22-90) */
    DXB_BEAM_CT_SPIRAL = 1,     /* dxmc::CTSpiralBeam<false>          R:src/libopendxmc/beamsettingsmodel.cpp:1158-1378 */
    DXB_BEAM_CT_SPIRAL_DUAL = 2,/* dxmc::CTSpiralDualEnergyBeam<false> R:...:1413-1832 */
    DXB_BEAM_CBCT = 3,          /* dxmc::CBCTBeam<false>              R:...:730-895 */
    DXB_BEAM_CT_SEQUENTIAL = 4, /* dxmc::CTSequentialBeam<false>      R:...:921-1131 */
    DXB_BEAM_PENCIL = 5,        /* dxmc::PencilBeam<false>            R:...:637-711 */
    DXB_BEAM_CTDI = 6           /* DXMClib-internal axial beam used by CT calibration */
} dxb_beam_type;

typedef struct dxb_spectrum { /* (E[], w[]) as produced by Tube — SURVEY.md §8a row a10 */
    uint32_t n;               /* n == 1: mono-energetic */
    const double* energy_kev; /* ascending, uniform step */
    const double* weight;     /* any positive scale */
} dxb_spectrum;

typedef struct dxb_bowtie {   /* dxmc::BowtieFilter(vector<pair<angle,weight>>) R:src/libopendxmc/bowtiefilterreader.cpp:74-93 */
    uint32_t n;               /* 0: no filter (weight 1) */
    const double* angle_rad;  /* unsorted, sign ignored */
    const double* weight;
} dxb_bowtie;

typedef struct dxb_aec {      /* dxmc::CTAECFilter(start, stop, weights) R:src/libopendxmc/datacontainer.cpp:37,59 */
    uint32_t n;               /* < 2: empty (weight 1) */
    double start[3], stop[3];
    const double* weights;
} dxb_aec;

typedef struct dxb_organ_aec { /* dxmc::CTOrganAECFilter R:src/libopendxmc/beamsettingsmodel.cpp:349-437 */
    int32_t use_filter;
    int32_t compensate_outside;
    double start_angle, stop_angle, ramp_angle; /* rad */
    double low_weight;
} dxb_organ_aec;

typedef struct dxb_beam_desc {
    int32_t  type;                    /* dxb_beam_type */
    int32_t  reserved0;
    uint64_t n_exposures;             /* DX, PENCIL: explicit; CT/CBCT: ignored (derived) */
    uint64_t particles_per_exposure;
    /* --- geometry --- */
    double position[3];               /* DX/PENCIL source position; CT_SEQUENTIAL/CTDI first slice centre */
    double direction[3];              /* PENCIL direction; CT_SEQUENTIAL scan normal; CBCT rotation axis */
    double cosines[2][3];             /* DX direction cosines */
    double half_angles[2];            /* DX, CBCT collimation half angles [rad] (stored value, see a2) */
    double start[3], stop[3];         /* CT_SPIRAL(_DUAL) */
    double isocenter[3];              /* CBCT */
    double sdd;                       /* source-detector distance */
    double fov;                       /* scan field of view (tube A) */
    double fov_b;                     /* tube B */
    double collimation;               /* total z collimation at isocentre */
    double pitch;
    double start_angle, step_angle;   /* rad */
    double stop_angle;                /* CBCT */
    uint64_t n_slices;                /* CT_SEQUENTIAL */
    double slice_spacing;
    double tube_b_offset_angle;       /* CT_SPIRAL_DUAL */
    double relative_mas_a, relative_mas_b;
    /* --- calibration --- */
    double ctdi;                      /* CTDIvol (spiral) or CTDIw (sequential) [mGy] */
    double ctdi_diameter;             /* phantom diameter [cm], default 32 */
    double dap;                       /* DX/CBCT dose-area product [mGy cm2] */
    double air_kerma;                 /* PENCIL [mGy] */
    double energy;                    /* PENCIL photon energy */
    /* --- per-tube data --- */
    dxb_spectrum spectrum[2];
    dxb_bowtie   bowtie[2];
    dxb_aec      aec;
    dxb_organ_aec organ_aec;
} dxb_beam_desc;

void dxb_beam_desc_init(dxb_beam_desc*, int type); /* zero + reference defaults for `type` */

/* Beam::numberOfExposures(), ::numberOfParticles(), ::exposure(i) — the geometry the GUI
 * reads at R:src/libopendxmc/beamactorcontainer.cpp:113-195 */
typedef struct dxb_exposure {
    double position[3];
    double cosines[2][3];
    double direction[3];      /* cross(cosines[0], cosines[1]) */
    double half_angles[2];
    double weight;            /* AEC * organ-AEC * relative mAs */
    uint64_t n_particles;
    int32_t tube;             /* 0 = A, 1 = B (selects spectrum/bowtie) */
    int32_t reserved;
} dxb_exposure;
uint64_t dxb_beam_number_of_exposures(const dxb_beam_desc*);
uint64_t dxb_beam_number_of_particles(const dxb_beam_desc*);
int dxb_beam_exposure(const dxb_beam_desc*, uint64_t index, dxb_exposure* out);
/* evaluate the (normalised) filters the way the kernels do */
double dxb_bowtie_weight(const dxb_bowtie*, double angle_rad);
double dxb_aec_weight(const dxb_aec*, const double position[3]);
double dxb_organ_aec_weight(const dxb_organ_aec*, double angle_rad);
double dxb_organ_aec_max_weight(const dxb_organ_aec*);
/* dose-per-history scale for non-CT beams: mGy per (keV/g) (DAP / air-kerma calibration) */
double dxb_beam_analytic_calibration(const dxb_beam_desc*);

/* ====================================================================== */
/* Progress (dxmc::TransportProgress) — R:src/libopendxmc/simulationpipeline.cpp:37,114-119,139,169,234,261 */
/* ====================================================================== */
typedef struct dxb_progress dxb_progress;
dxb_progress* dxb_progress_create(void);
void dxb_progress_destroy(dxb_progress*);
void dxb_progress_read(const dxb_progress*, uint64_t* done, uint64_t* total);
void dxb_progress_stop(dxb_progress*);            /* setStopSimulation() */
int  dxb_progress_continue(const dxb_progress*);  /* continueSimulation() */
void dxb_progress_reset(dxb_progress*);
/* message(): copies a NUL-terminated text ("hh:mm:ss remaining") into buf */
int  dxb_progress_message(const dxb_progress*, char* buf, int cap);

/* ====================================================================== */
/* Context = World<AAVoxelGrid<5,C,255>> + Transport on one or more GPUs    */
/* ====================================================================== */
typedef struct dxb_ctx dxb_ctx;

/* cuda_devices == NULL && n_devices == 0: use the current device.
 * Several devices (one process drives them all, like the reference's single worker,
 * R:src/libopendxmc/simulationpipeline.cpp:161-167): the histories of every beam are sharded over the devices and the
 * library exchanges the tallies itself - dxb_set_grid uploads one slab of the caller's arrays per device and
 * all-gathers the packed voxels over NVLink, the per-beam tallies are double-buffered, every device pulls its voxel slab
 * of the peers' tallies with the copy engines underneath the NEXT beam's transport kernels and converts it to dose, the
 * dose score stays distributed and is read out by all devices concurrently (opendxmc_b200/csrc/exchange.cu).  Same
 * C ABI, same results (integer tallies: bitwise identical for any device count). */
int  dxb_create(dxb_ctx** out, const int* cuda_devices, int n_devices);
void dxb_destroy(dxb_ctx*);
const char* dxb_last_error(const dxb_ctx*); /* text of the last failure on this context */

/* setData for one process per GPU (every rank needs the full replica).  Instead of N full uploads,
 * rank r uploads and packs only voxels [voxel_begin, voxel_end) of the caller's FULL-size arrays; the ranks then
 * exchange the packed 4-byte slabs (dxb_grid_buffers: `voxels`, u32 x n_voxels) and take the element-wise maximum of
 * `max_density_bits` (u32 x 257: per-material density maxima, [256] = largest material index) over NVLink - one
 * broadcast per slab + one all-reduce(MAX), opendxmc_b200/distributed.py: set_grid_sharded - and call dxb_finish_grid
 * (majorant, material-index check).  Same result as dxb_set_grid on every rank, 1/N of the host-to-device bytes each. */
int dxb_set_grid_sharded(dxb_ctx*, const uint64_t dim[3], const double spacing_cm[3], const double* density, const uint8_t* material,
                         uint64_t voxel_begin, uint64_t voxel_end);
int dxb_grid_buffers(dxb_ctx*, void** voxels, void** max_density_bits, uint64_t* n_voxels);
int dxb_finish_grid(dxb_ctx*);

/* AAVoxelGrid::setData(materials part) — R:src/libopendxmc/simulationpipeline.cpp:134-149 */
int dxb_set_materials(dxb_ctx*, uint32_t n, const dxb_material* const* materials);
/* AAVoxelGrid::setData(dims, density, material) + setSpacing + World::build():
 * index order x fastest: i + j*nx + k*nx*ny (R:src/libopendxmc/otherphantomimportpipeline.cpp:44);
 * grid centred on the origin (R:src/libopendxmc/datacontainer.cpp:174-178).
 * Copies to the device, packs voxels, builds the per-energy majorant. Clears all dose. */
int dxb_set_grid(dxb_ctx*, const uint64_t dim[3], const double spacing_cm[3],
                 const double* density, const uint8_t* material);
int dxb_set_grid_center(dxb_ctx*, const double center_cm[3]); /* default origin */

/* Transport knobs */
/* Philox keys.  dxb_set_seed sets the BASE key (default 0x0DDC0FFEE) and restarts the context's beam counter; the
 * k-th dxb_run / dxb_run_transport after it runs on key  base + k * DXB_BEAM_KEY_STRIDE  (k = 0: the base key itself), so
 * that the beams of one job (R:src/libopendxmc/simulationpipeline.cpp:161-167 calls transport() once per beam) draw
 * independent streams; the nested CTDI calibration run of a beam uses  beam key ^ DXB_CALIBRATION_KEY_XOR.  Every rank
 * of a sharded job counts the same beams, so results stay independent of the GPU count.  dxb_last_beam_key returns
 * the key of the last beam (what a test hands to the CPU oracle). */
#define DXB_BEAM_KEY_STRIDE 0x9E3779B97F4A7C15ull
#define DXB_CALIBRATION_KEY_XOR 0x4354444900000000ull /* "CTDI" in the high word */
int dxb_set_seed(dxb_ctx*, uint64_t seed);
uint64_t dxb_last_beam_key(const dxb_ctx*);
int dxb_set_history_range(dxb_ctx*, uint64_t rank, uint64_t world); /* this context runs block `rank` of `world` of every exposure (multi-process sharding) */
/* the sharding rule itself (host arithmetic, no device needed): histories are dealt to shards in blocks of
 * DXB_SHARD_BLOCK consecutive ids, round-robin.  dxb_shard_local_count = local indices owned by `rank` (padded to
 * whole blocks); dxb_shard_history_id maps a local index to the global history id (ids >= n_total are skipped). */
#define DXB_SHARD_BLOCK 65536u
uint64_t dxb_shard_local_count(uint64_t n_total, uint64_t rank, uint64_t world);
uint64_t dxb_shard_history_id(uint64_t local_index, uint64_t rank, uint64_t world);
int dxb_set_calibration_histories(dxb_ctx*, uint64_t n);      /* nested CTDI run size */
int dxb_set_stream(dxb_ctx*, void* cuda_stream);              /* launch on a caller-owned stream (device 0 of the ctx) */
/* Tuning knobs; none changes a result except the four that select the tracking variant - "dense_box" and
 * "local_majorant" (-1 auto, 0 off, 1 on), "dense_theta", "slab_cm" - which change the random walk, not its expectation
 * (0 / 0 = plain Woodcock tracking with one majorant, the reference's rule).  Unknown keys return DXB_EINVAL. */
int dxb_set_option(dxb_ctx*, const char* key, double value);

/* Transport::operator()(world, beam, progress, useBeamCalibration) — R:src/libopendxmc/simulationpipeline.cpp:165.
 * Blocking. Clears the energy tallies, runs all histories of the beam, reduces across the
 * context's devices, converts energy to dose and ACCUMULATES into the dose score, exactly like
 * repeated transport() calls on one world. */
int dxb_run(dxb_ctx*, const dxb_beam_desc*, int physics_mode, int use_beam_calibration, dxb_progress*);

/* Split form of dxb_run for multi-process sharding (one process per GPU):
 *   dxb_run_transport  — tallies only (no energy->dose), may be called on every rank;
 *   the caller sums the tally buffer across ranks (NCCL, int64 sum) in place;
 *   dxb_finish_beam    — calibration factor + energy->dose on the summed tallies. */
int dxb_run_transport(dxb_ctx*, const dxb_beam_desc*, int physics_mode, dxb_progress*);
int dxb_finish_beam(dxb_ctx*, const dxb_beam_desc*, int physics_mode, int use_beam_calibration, double* factor_out);
/* device pointer + element count (int64 words) of the raw fixed-point tally buffer of device 0 */
int dxb_tally_buffer(dxb_ctx*, void** device_ptr, uint64_t* n_words);

/* Fused exchange + finish for one process per GPU inside one NVSwitch domain (replaces "reduce to rank 0, then
 * dxb_finish_beam"; same arithmetic as the tail of dxmc::Transport::operator(), call site
 * R:src/libopendxmc/simulationpipeline.cpp:165).
 *   dxb_set_tally_storage  — make caller-provided device memory (4 x u64 per voxel, 32-byte aligned; e.g. a
 *                            symmetric allocation that is peer- and multicast-mapped on every rank) the tally
 *                            buffer of this context; NULL returns to library-owned storage.  The library zeroes it.
 *   dxb_finish_beam_sharded — this rank converts voxels [voxel_begin, voxel_end) of the SUM over ranks of the
 *                            tallies into dose/variance/events, reading the sum in ONE kernel either through
 *                            `multicast_tally` (NVSwitch in-flight add, multimem.ld_reduce) or, when that is NULL,
 *                            from its own buffer plus the `n_peers` peer-mapped buffers `peer_tallies` (NVLink
 *                            P2P loads).  The caller separates it from the transport of all ranks, and from the
 *                            next beam, by a barrier.  The dose score of a slab lives on the rank that owns it
 *                            until the caller gathers it (dxb_dose_buffers gives the device arrays).
 *   Every rank computes the beam calibration factor itself (deterministic), so none is exchanged. */
int dxb_set_tally_storage(dxb_ctx*, void* device_ptr, uint64_t n_words);
int dxb_finish_beam_sharded(dxb_ctx*, const dxb_beam_desc*, int physics_mode, int use_beam_calibration,
                            const void* multicast_tally, const void* const* peer_tallies, int n_peers,
                            uint64_t voxel_begin, uint64_t voxel_end, double* factor_out);
/* The library-managed exchange for ONE PROCESS PER GPU (torchrun): the same pipelined exchange a multi-device context
 * runs, with the peers' tally buffers mapped through CUDA IPC instead of living in the same process.
 *   dxb_exchange_export  - after dxb_set_grid*: allocates the second tally buffer and writes DXB_EXCHANGE_HANDLE_BYTES of
 *                          opaque handle data; the caller all-gathers them over the ranks (any transport: they are bytes);
 *   dxb_exchange_import  - maps every peer's buffers, sets the history shard (rank, world) and turns the context into an
 *                          exchanging one: dxb_finish_beam then enqueues (a) the clear of the previous beam's buffer,
 *                          (b) the copy-engine pulls of this rank's voxel slab from every peer, (c) slab reduce -> dose,
 *                          and returns; dxb_get_dose_range / dxb_flush wait for it.  Per beam the caller runs
 *                              dxb_run_transport;  barrier over the ranks;  dxb_finish_beam
 *                          (the barrier is the only synchronisation the library cannot do itself across processes).
 *                          The nested CTDI calibration run of a beam is sharded over the ranks too: every rank runs its
 *                          share of the calibration histories and publishes five integer sums in a small mailbox the
 *                          peers read (so dxb_finish_beam with use_beam_calibration waits for all ranks to get there).
 *   dxb_flush            - waits until every enqueued transport / exchange of the context has completed.
 *   dxb_exchange_times   - device time of the last flushed exchange on device 0: {pulls, slab reduce -> dose, clear} [ms]. */
#define DXB_EXCHANGE_HANDLE_BYTES 192
int dxb_exchange_export(dxb_ctx*, void* handles);
int dxb_exchange_import(dxb_ctx*, uint64_t rank, uint64_t world, const void* all_handles /* world x DXB_EXCHANGE_HANDLE_BYTES */);
int dxb_exchange_close(dxb_ctx*); /* unmaps the peers' buffers (after a flush and a barrier, before any rank destroys its context) */
int dxb_flush(dxb_ctx*);
int dxb_exchange_times(const dxb_ctx*, double out_ms[3]);
/* Device-side stopwatch over ALL devices and streams of the context (bench.py): begin flushes and records a CUDA event
 * on every device, end records after everything enqueued since and returns the largest elapsed time. */
int dxb_timer_begin(dxb_ctx*);
int dxb_timer_end(dxb_ctx*, double* ms_max);

/* device arrays of the accumulated dose score of device 0: f64 dose [mGy], f64 variance, u64 events, n_voxels each */
int dxb_dose_buffers(dxb_ctx*, void** dose, void** variance, void** n_events, uint64_t* n_voxels);

/* doseScored(i).dose()/variance()/numberOfEvents() for all i — R:src/libopendxmc/simulationpipeline.cpp:174-219.
 * Any pointer may be NULL. dose [mGy], variance [mGy^2]. */
int dxb_get_dose(dxb_ctx*, double* dose, double* variance, uint64_t* n_events);
/* The same read-out for voxels [voxel_begin, voxel_end) only: element i of the FULL-size caller arrays is written for
 * voxel_begin <= i < voxel_end, nothing else is touched.  With one process per GPU and the fused exchange every rank
 * holds the dose score of its own slab (dxb_finish_beam_sharded); the ranks then read their slabs out in parallel,
 * each over its own PCIe link, into arrays the caller placed in host memory shared by the processes - instead of
 * gathering 24 B/voxel to rank 0 and pushing all of it through one link. */
int dxb_get_dose_range(dxb_ctx*, uint64_t voxel_begin, uint64_t voxel_end, double* dose, double* variance, uint64_t* n_events);
/* the per-beam energy tallies of the LAST beam (keV, keV^2, count) */
int dxb_get_energy_scored(dxb_ctx*, double* energy, double* energy_sq, uint64_t* n_events);
int dxb_clear_dose(dxb_ctx*);

/* The reference's post-processing at R:src/libopendxmc/simulationpipeline.cpp:174-232 done on the
 * device before the copy-out: air mask (material==0 -> 0) and the uGy rescale.
 * units_out receives "mGy" or "uGy". */
int dxb_get_dose_postprocessed(dxb_ctx*, int delete_air_dose, double* dose, double* variance,
                               double* n_events, char units_out[4]);

/* Per-organ mass-weighted mean dose, R:src/libopendxmc/dosetablepipeline.cpp:60-84. */
int dxb_organ_dose(dxb_ctx*, const uint8_t* organ, uint32_t n_organs,
                   double* dose_out, double* mass_out, uint64_t* n_voxels_out, double* variance_out);

/* Run statistics of the last dxb_run / dxb_run_transport on this context. */
typedef struct dxb_run_stats {
    uint64_t histories;       /* histories simulated by this context */
    uint64_t steps;           /* tentative Woodcock steps (voxel fetches) */
    uint64_t interactions;    /* accepted (real) interactions */
    uint64_t deposits;        /* energy-depositing events (tally updates) */
    uint64_t kernel_launches; /* transport kernel launches */
    double   transport_ms;    /* device time of the transport kernels (CUDA events) */
    double   total_ms;        /* device time incl. reduce + energy->dose */
    double   calibration_ms;  /* nested CTDI run */
    double   calibration_factor;
    double   energy_emitted_kev;   /* sum of E*w over sampled histories */
    double   energy_deposited_kev; /* sum over tallies */
    uint64_t hops;            /* slab-local majorants: tentative steps that ended on a slab face (no voxel fetch) */
    int32_t  local_majorant;  /* 1: the last beam ran on the slab-local majorant build of the kernel */
    int32_t  dense_box;       /* 1: the last beam ran on the dense-box build of the kernel (hops = box entries then) */
    uint64_t voxel_fetches;   /* voxel gathers actually issued (<= steps: the brick pre-filter skips certainly-virtual collisions) */
} dxb_run_stats;
int dxb_get_run_stats(const dxb_ctx*, dxb_run_stats* out);

/* Device-side table lookups, for the "lookups within 1e-6" parity test: evaluates
 * {photo, incoh, coh, total}*1 [cm2/g] of material `m` at n energies with the kernels' own
 * float code path. */
int dxb_device_attenuation(dxb_ctx*, uint32_t material_index, int physics_mode,
                           const double* energy_kev, uint32_t n, float* out4);
/* majorant mu_max(E) [1/cm] as the kernels interpolate it */
int dxb_device_majorant(dxb_ctx*, const double* energy_kev, uint32_t n, float* out);
/* The slab-local majorant table built with the grid (option "local_majorant": -1 auto, 0 off, 1 on; "slab_cm": target slab
 * thickness): slabs of 2^shift voxel layers along z; inside slab s and energy band b (= energy node index >> 5) the kernels
 * track with mu_max(E) * ratio[s * 16 + b], ratio in (0, 1].  *useful = 1 when auto mode would use it.  ratio may be NULL
 * (sizes only); it receives n_slabs * 16 floats.  This is what a test hands to the CPU oracle so that both track identically. */
int dxb_get_local_majorant(dxb_ctx*, int* n_slabs, int* shift, int* useful, float* ratio);
/* The dense box built with the grid (option "dense_box": -1 auto, 0 off, 1 on; "dense_theta": a voxel is thin when its
 * attenuation stays below this fraction of the majorant at every energy, default 0.02): box = first voxel index x y z and one
 * past the last x y z of the bounding box of all voxels that are not thin, faces = the same as coordinates [cm] (f32, as the
 * kernel uses them).  Inside the box the kernels track with mu_max(E), in the rest of the grid with mu_max(E) * ratio[band]
 * (band = energy node index >> 5) - the air around a patient then costs one flight instead of a tentative step every ~2 cm.
 * *built = 0: no box (all voxels thin, or the option was 0 when the grid was set); *useful = 1 when auto mode uses it.  Any
 * pointer may be NULL.  This is what a test hands to the CPU oracle so that both track identically. */
int dxb_get_dense_box(dxb_ctx*, int* built, int* useful, int box[6], float faces[6], float ratio[16]);

/* CT segmentation, SURVEY §8f-1: HU -> (material, density)
 * R:src/libopendxmc/ctsegmentationpipeline.cpp:113-169 */
int dxb_segment_ct(dxb_ctx*, const double* hu, uint64_t n, const dxb_tube_desc* tube,
                   uint8_t* material_out, double* density_out,
                   dxb_material** materials_out /* [5], caller destroys */);

/* ICRP 110 / 143 voxel-phantom import, SURVEY §8f-2: what ICRPPhantomImportPipeline::importPhantom does with the
 * phantom's organ array and its two text tables (R:src/libopendxmc/icrpphantomimportpipeline.cpp:209-351; parsers :59-205):
 * "Air" appended as organ 0 / medium 0; with remove_arms the organs named *arm*, *hand*, *Humeri*, *Ulnae* become air;
 * organs absent from the array are dropped and the rest renumbered consecutively, unused media likewise; material and
 * density arrays by organ lookup.  The tables are passed as the TEXT of `<phantom>_organs.dat` / `<phantom>_media.dat`.
 *   dxb_icrp_import  device path: presence scan + one look-up-table gather per voxel over the organ array (n = nx*ny*nz
 *                    bytes, x fastest), writes organ / material (u8) and density (f64) arrays of n elements;
 *   dxb_icrp_plan    the host-side rules alone, given which organ values occur (`present[v] != 0`); dxb_icrp_luts returns
 *                    the three 256-entry tables the device pass applies (organ value -> organ index, medium, density).
 * The plan lists the surviving organs (name, density, medium index) and media (name, composition in mass %, zero
 * entries kept like the reference, to hand to dxb_material_by_weight). */
typedef struct dxb_icrp dxb_icrp;
int dxb_icrp_import(dxb_ctx*, const uint8_t* organ_in, uint64_t n, const char* organs_dat, const char* media_dat, int remove_arms,
                    uint8_t* organ_out, uint8_t* material_out, double* density_out, dxb_icrp** plan_out);
int dxb_icrp_plan(dxb_icrp** out, const char* organs_dat, const char* media_dat, int remove_arms, const uint8_t present[256]);
int dxb_icrp_luts(const dxb_icrp*, uint8_t organ_lut[256], uint8_t material_lut[256], double density_lut[256]);
void dxb_icrp_destroy(dxb_icrp*);
uint32_t dxb_icrp_n_organs(const dxb_icrp*);
const char* dxb_icrp_organ_name(const dxb_icrp*, uint32_t index);
double dxb_icrp_organ_density(const dxb_icrp*, uint32_t index);
uint32_t dxb_icrp_organ_medium(const dxb_icrp*, uint32_t index);
uint32_t dxb_icrp_n_media(const dxb_icrp*);
const char* dxb_icrp_medium_name(const dxb_icrp*, uint32_t index);
int dxb_icrp_medium_composition(const dxb_icrp*, uint32_t index, uint32_t* Z, double* weight, int cap); /* returns the element count */

/* ====================================================================== */
/* HDF5 save files, SURVEY §8f-3 (R:src/libopendxmc/hdf5wrapper.cpp)        */
/* ====================================================================== */
/* A minimal reader / writer for the subset of the HDF5 format OpenDXMC's save files use (opendxmc_b200/csrc/h5mini.*:
 * superblock version 0, old-style groups, version-1 object headers, attributes, f64 / u64 / u8 / variable-length string
 * datasets stored contiguously or as one deflate-6 chunk; the reader also walks multi-chunk B-trees).  The image has no
 * HDF5 library: the reader is pinned on a genuine libhdf5-written file, the writer on the reader, and the reference's own
 * hdf5wrapper.cpp round-trips through both (tests/stubs/H5Cpp.h is a shim over this API).
 * Object-level API: an in-memory image of a file is built (create + put_*) and written by dxb_h5_save, or parsed by
 * dxb_h5_open and queried.  Paths are "/group/sub/name"; dims are in HDF5 order (slowest first). */
typedef struct dxb_h5 dxb_h5;
enum { DXB_H5_UNKNOWN = 0, DXB_H5_F64 = 1, DXB_H5_U64 = 2, DXB_H5_U8 = 3, DXB_H5_STRING = 4, DXB_H5_I64 = 5, DXB_H5_I32 = 6,
       DXB_H5_U32 = 7, DXB_H5_F32 = 8, DXB_H5_U16 = 9, DXB_H5_I16 = 10, DXB_H5_I8 = 11 };
dxb_h5* dxb_h5_create(void);
int dxb_h5_open(dxb_h5** out, const char* path);
void dxb_h5_close(dxb_h5*);
const char* dxb_h5_error(const dxb_h5*);
int dxb_h5_save(dxb_h5*, const char* path);
int dxb_h5_exists(const dxb_h5*, const char* path);               /* 0 no, 1 group, 2 dataset (H5File::nameExists) */
int dxb_h5_make_group(dxb_h5*, const char* path);                 /* creates missing parents */
const char* dxb_h5_list(dxb_h5*, const char* group_path);         /* "g name\n" / "d name\n" / "a name\n" lines; valid until the next call */
int dxb_h5_put_dataset(dxb_h5*, const char* path, int type, int rank, const uint64_t* dims, const void* data, int deflate);
int dxb_h5_put_strings(dxb_h5*, const char* path, uint64_t n, const char* const* strings);
int dxb_h5_put_attribute(dxb_h5*, const char* group_path, const char* name, int type, int64_t n /* < 0: scalar */, const void* data);
int dxb_h5_dataset_info(const dxb_h5*, const char* path, int* type, int* rank, uint64_t dims[8], int* deflate);
int dxb_h5_dataset_read(const dxb_h5*, const char* path, void* out, uint64_t out_bytes);
const char* dxb_h5_dataset_string(const dxb_h5*, const char* path, uint64_t index);
int dxb_h5_attribute_info(const dxb_h5*, const char* group_path, const char* name, int* type, int64_t* n /* -1: scalar */);
int dxb_h5_attribute_read(const dxb_h5*, const char* group_path, const char* name, void* out, uint64_t out_bytes);
const char* dxb_h5_attribute_string(const dxb_h5*, const char* group_path, const char* name, uint64_t index);

/* Scene-level: drive the engine head-less from a file the GUI wrote, and hand the result back in the same format.
 *   dxb_load_scene  reads "dimensions", "spacing", "densityarray", "materialarray", "materialnames", "materialcomposition"
 *                   (R:src/libopendxmc/hdf5wrapper.cpp:1070-1118), builds the materials like the worker does
 *                   (parseCompoundStr -> Material::byWeight, R:...simulationpipeline.cpp:136) and calls dxb_set_materials +
 *                   dxb_set_grid; the file's other content (organ array, names, AEC data) stays with the context;
 *   dxb_save_dose   writes everything dxb_load_scene read plus "dosearray", "dosevariancearray", "doseeventcountarray"
 *                   (f64, HDF5 dims (nz, ny, nx), one deflate-6 chunk each, :121-151, :436-447) after the driver's
 *                   post-processing (air mask if delete_air_dose, uGy rule; units_out = "mGy" / "uGy"). */
int dxb_load_scene(dxb_ctx*, const char* path, uint64_t dim_out[3], double spacing_cm_out[3], uint32_t* n_materials_out);
int dxb_save_dose(dxb_ctx*, const char* path, int delete_air_dose, char units_out[4]);

int dxb_abi_version(void);
/* identity of the transport kernel compiled into this library: hash of its sources and compiler flags (profiles taken
 * from another build do not describe it; bench.py compares it with the id stored next to an ncu traffic capture) */
const char* dxb_kernel_build_id(void);
int dxb_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DXB_H */

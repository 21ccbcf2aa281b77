"""ctypes declarations for include/dxb.h — the C ABI of libdxmc_b200.so.

This module only *declares*; it never computes.  If the shared library has not been built
(`make` or `python -c "import __graft_entry__ as g; g.build()"`) importing it raises: the product
has no CPU fallback (include/dxb.h, "Rules of the ABI").
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdxmc_b200.so")

DXB_EXCHANGE_HANDLE_BYTES = 192
DXB_OK, DXB_EINVAL, DXB_EMATERIAL, DXB_ECUDA, DXB_ESTATE, DXB_ECANCELLED, DXB_ENOMEM = range(7)
STATUS_NAMES = ["DXB_OK", "DXB_EINVAL", "DXB_EMATERIAL", "DXB_ECUDA", "DXB_ESTATE", "DXB_ECANCELLED", "DXB_ENOMEM"]

PHYSICS_NONE, PHYSICS_LIVERMORE, PHYSICS_IA = 0, 1, 2
BEAM_DX, BEAM_CT_SPIRAL, BEAM_CT_SPIRAL_DUAL, BEAM_CBCT, BEAM_CT_SEQUENTIAL, BEAM_PENCIL, BEAM_CTDI = range(7)
MAX_SHELLS = 5
TUBE_MAX_FILT = 8

c_double_p = C.POINTER(C.c_double)
c_u8_p = C.POINTER(C.c_uint8)
c_u32_p = C.POINTER(C.c_uint32)
c_u64_p = C.POINTER(C.c_uint64)
c_float_p = C.POINTER(C.c_float)


class dxb_shell(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "binding_energy_kev", "n_electrons", "n_electrons_fraction", "photo_fraction_above",
        "fluor_yield", "fluor_energy_kev", "compton_j0")]


class dxb_material_tables(C.Structure):
    _fields_ = [
        ("n_energy", C.c_uint32), ("e_min_kev", C.c_double), ("e_max_kev", C.c_double),
        ("photo", c_double_p), ("incoh", c_double_p), ("coh", c_double_p), ("incoh_kn", c_double_p),
        ("coh_thomson", c_double_p), ("etr", c_double_p),
        ("n_x", C.c_uint32), ("x_min", C.c_double), ("x_max", C.c_double),
        ("ff_cdf", c_double_p), ("sf", c_double_p),
        ("n_shells", C.c_uint32), ("shells", dxb_shell * MAX_SHELLS),
        ("rest_electrons_fraction", C.c_double), ("rest_compton_j0", C.c_double), ("electrons_per_gram", C.c_double), ("effective_z", C.c_double),
        ("nodes_per_octave_e", C.c_uint32), ("nodes_per_octave_x", C.c_uint32),
    ]


class dxb_tube_desc(C.Structure):
    _fields_ = [
        ("voltage_kv", C.c_double), ("anode_angle_deg", C.c_double), ("energy_resolution_kev", C.c_double),
        ("n_filt", C.c_uint32), ("filt_Z", C.c_uint32 * TUBE_MAX_FILT), ("filt_mm", C.c_double * TUBE_MAX_FILT),
    ]


class dxb_spectrum(C.Structure):
    _fields_ = [("n", C.c_uint32), ("energy_kev", c_double_p), ("weight", c_double_p)]


class dxb_bowtie(C.Structure):
    _fields_ = [("n", C.c_uint32), ("angle_rad", c_double_p), ("weight", c_double_p)]


class dxb_aec(C.Structure):
    _fields_ = [("n", C.c_uint32), ("start", C.c_double * 3), ("stop", C.c_double * 3), ("weights", c_double_p)]


class dxb_organ_aec(C.Structure):
    _fields_ = [
        ("use_filter", C.c_int32), ("compensate_outside", C.c_int32),
        ("start_angle", C.c_double), ("stop_angle", C.c_double), ("ramp_angle", C.c_double), ("low_weight", C.c_double),
    ]


class dxb_beam_desc(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("reserved0", C.c_int32),
        ("n_exposures", C.c_uint64), ("particles_per_exposure", C.c_uint64),
        ("position", C.c_double * 3), ("direction", C.c_double * 3), ("cosines", (C.c_double * 3) * 2),
        ("half_angles", C.c_double * 2), ("start", C.c_double * 3), ("stop", C.c_double * 3), ("isocenter", C.c_double * 3),
        ("sdd", C.c_double), ("fov", C.c_double), ("fov_b", C.c_double), ("collimation", C.c_double), ("pitch", C.c_double),
        ("start_angle", C.c_double), ("step_angle", C.c_double), ("stop_angle", C.c_double),
        ("n_slices", C.c_uint64), ("slice_spacing", C.c_double), ("tube_b_offset_angle", C.c_double),
        ("relative_mas_a", C.c_double), ("relative_mas_b", C.c_double),
        ("ctdi", C.c_double), ("ctdi_diameter", C.c_double), ("dap", C.c_double), ("air_kerma", C.c_double), ("energy", C.c_double),
        ("spectrum", dxb_spectrum * 2), ("bowtie", dxb_bowtie * 2), ("aec", dxb_aec), ("organ_aec", dxb_organ_aec),
    ]


class dxb_exposure(C.Structure):
    _fields_ = [
        ("position", C.c_double * 3), ("cosines", (C.c_double * 3) * 2), ("direction", C.c_double * 3),
        ("half_angles", C.c_double * 2), ("weight", C.c_double), ("n_particles", C.c_uint64),
        ("tube", C.c_int32), ("reserved", C.c_int32),
    ]


class dxb_run_stats(C.Structure):
    _fields_ = [
        ("histories", C.c_uint64), ("steps", C.c_uint64), ("interactions", C.c_uint64), ("deposits", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("transport_ms", C.c_double), ("total_ms", C.c_double), ("calibration_ms", C.c_double),
        ("calibration_factor", C.c_double), ("energy_emitted_kev", C.c_double), ("energy_deposited_kev", C.c_double),
        ("hops", C.c_uint64), ("local_majorant", C.c_int32), ("dense_box", C.c_int32), ("voxel_fetches", C.c_uint64),
    ]


VP = C.c_void_p
# name -> (restype, argtypes); this table is also what tests/test_capi_symbols.py checks against include/dxb.h
SIGNATURES = {
    "dxb_abi_version": (C.c_int, []),
    "dxb_kernel_build_id": (C.c_char_p, []),
    "dxb_device_count": (C.c_int, []),
    "dxb_material_by_weight": (C.c_int, [C.POINTER(VP), C.c_uint32, c_u32_p, c_double_p]),
    "dxb_material_by_nist_name": (C.c_int, [C.POINTER(VP), C.c_char_p]),
    "dxb_material_by_chemical_formula": (C.c_int, [C.POINTER(VP), C.c_char_p]),
    "dxb_material_destroy": (None, [VP]),
    "dxb_material_attenuation": (C.c_int, [VP, C.c_double, c_double_p]),
    "dxb_material_mass_energy_transfer": (C.c_double, [VP, C.c_double]),
    "dxb_material_effective_z": (C.c_double, [VP]),
    "dxb_material_form_factor": (C.c_double, [VP, C.c_double]),
    "dxb_material_scatter_factor": (C.c_double, [VP, C.c_double]),
    "dxb_material_tables_get": (C.c_int, [VP, C.POINTER(dxb_material_tables)]),
    "dxb_material_from_tables": (C.c_int, [C.POINTER(VP), C.POINTER(dxb_material_tables)]),
    "dxb_table_n_energy": (C.c_uint32, []),
    "dxb_table_e_min": (C.c_double, []),
    "dxb_table_e_max": (C.c_double, []),
    "dxb_nist_count": (C.c_int, []),
    "dxb_nist_name": (C.c_char_p, [C.c_int]),
    "dxb_nist_density": (C.c_double, [C.c_char_p]),
    "dxb_nist_composition": (C.c_int, [C.c_char_p, c_u32_p, c_double_p, C.c_int]),
    "dxb_atom_symbol": (C.c_char_p, [C.c_uint32]),
    "dxb_atom_weight": (C.c_double, [C.c_uint32]),
    "dxb_atom_standard_density": (C.c_double, [C.c_uint32]),
    "dxb_tube_energies": (C.c_int, [C.POINTER(dxb_tube_desc), c_double_p, C.c_int]),
    "dxb_tube_spectrum": (C.c_int, [C.POINTER(dxb_tube_desc), c_double_p, C.c_int, C.c_int, c_double_p]),
    "dxb_tube_mean_energy": (C.c_double, [C.POINTER(dxb_tube_desc)]),
    "dxb_tube_al_half_value_layer_mm": (C.c_double, [C.POINTER(dxb_tube_desc)]),
    "dxb_beam_desc_init": (None, [C.POINTER(dxb_beam_desc), C.c_int]),
    "dxb_beam_number_of_exposures": (C.c_uint64, [C.POINTER(dxb_beam_desc)]),
    "dxb_beam_number_of_particles": (C.c_uint64, [C.POINTER(dxb_beam_desc)]),
    "dxb_beam_exposure": (C.c_int, [C.POINTER(dxb_beam_desc), C.c_uint64, C.POINTER(dxb_exposure)]),
    "dxb_bowtie_weight": (C.c_double, [C.POINTER(dxb_bowtie), C.c_double]),
    "dxb_aec_weight": (C.c_double, [C.POINTER(dxb_aec), c_double_p]),
    "dxb_organ_aec_weight": (C.c_double, [C.POINTER(dxb_organ_aec), C.c_double]),
    "dxb_organ_aec_max_weight": (C.c_double, [C.POINTER(dxb_organ_aec)]),
    "dxb_beam_analytic_calibration": (C.c_double, [C.POINTER(dxb_beam_desc)]),
    "dxb_progress_create": (VP, []),
    "dxb_progress_destroy": (None, [VP]),
    "dxb_progress_read": (None, [VP, c_u64_p, c_u64_p]),
    "dxb_progress_stop": (None, [VP]),
    "dxb_progress_continue": (C.c_int, [VP]),
    "dxb_progress_reset": (None, [VP]),
    "dxb_progress_message": (C.c_int, [VP, C.c_char_p, C.c_int]),
    "dxb_create": (C.c_int, [C.POINTER(VP), C.POINTER(C.c_int), C.c_int]),
    "dxb_destroy": (None, [VP]),
    "dxb_last_error": (C.c_char_p, [VP]),
    "dxb_set_materials": (C.c_int, [VP, C.c_uint32, C.POINTER(VP)]),
    "dxb_set_grid": (C.c_int, [VP, c_u64_p, c_double_p, c_double_p, c_u8_p]),
    "dxb_set_grid_center": (C.c_int, [VP, c_double_p]),
    "dxb_set_seed": (C.c_int, [VP, C.c_uint64]),
    "dxb_last_beam_key": (C.c_uint64, [VP]),
    "dxb_exchange_export": (C.c_int, [VP, VP]),
    "dxb_exchange_import": (C.c_int, [VP, C.c_uint64, C.c_uint64, VP]),
    "dxb_exchange_close": (C.c_int, [VP]),
    "dxb_flush": (C.c_int, [VP]),
    "dxb_exchange_times": (C.c_int, [VP, c_double_p]),
    "dxb_timer_begin": (C.c_int, [VP]),
    "dxb_timer_end": (C.c_int, [VP, c_double_p]),
    "dxb_set_history_range": (C.c_int, [VP, C.c_uint64, C.c_uint64]),
    "dxb_shard_local_count": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint64]),
    "dxb_shard_history_id": (C.c_uint64, [C.c_uint64, C.c_uint64, C.c_uint64]),
    "dxb_set_calibration_histories": (C.c_int, [VP, C.c_uint64]),
    "dxb_set_stream": (C.c_int, [VP, VP]),
    "dxb_set_option": (C.c_int, [VP, C.c_char_p, C.c_double]),
    "dxb_run": (C.c_int, [VP, C.POINTER(dxb_beam_desc), C.c_int, C.c_int, VP]),
    "dxb_run_transport": (C.c_int, [VP, C.POINTER(dxb_beam_desc), C.c_int, VP]),
    "dxb_finish_beam": (C.c_int, [VP, C.POINTER(dxb_beam_desc), C.c_int, C.c_int, c_double_p]),
    "dxb_tally_buffer": (C.c_int, [VP, C.POINTER(VP), c_u64_p]),
    "dxb_set_tally_storage": (C.c_int, [VP, VP, C.c_uint64]),
    "dxb_finish_beam_sharded": (C.c_int, [VP, C.POINTER(dxb_beam_desc), C.c_int, C.c_int, VP, C.POINTER(VP), C.c_int, C.c_uint64,
                                          C.c_uint64, C.POINTER(C.c_double)]),
    "dxb_dose_buffers": (C.c_int, [VP, C.POINTER(VP), C.POINTER(VP), C.POINTER(VP), c_u64_p]),
    "dxb_get_dose": (C.c_int, [VP, c_double_p, c_double_p, c_u64_p]),
    "dxb_set_grid_sharded": (C.c_int, [VP, C.POINTER(C.c_uint64), c_double_p, c_double_p, c_u8_p, C.c_uint64, C.c_uint64]),
    "dxb_grid_buffers": (C.c_int, [VP, C.POINTER(VP), C.POINTER(VP), c_u64_p]),
    "dxb_finish_grid": (C.c_int, [VP]),
    "dxb_get_dose_range": (C.c_int, [VP, C.c_uint64, C.c_uint64, c_double_p, c_double_p, c_u64_p]),
    "dxb_get_energy_scored": (C.c_int, [VP, c_double_p, c_double_p, c_u64_p]),
    "dxb_clear_dose": (C.c_int, [VP]),
    "dxb_get_dose_postprocessed": (C.c_int, [VP, C.c_int, c_double_p, c_double_p, c_double_p, C.c_char_p]),
    "dxb_organ_dose": (C.c_int, [VP, c_u8_p, C.c_uint32, c_double_p, c_double_p, c_u64_p, c_double_p]),
    "dxb_get_run_stats": (C.c_int, [VP, C.POINTER(dxb_run_stats)]),
    "dxb_device_attenuation": (C.c_int, [VP, C.c_uint32, C.c_int, c_double_p, C.c_uint32, c_float_p]),
    "dxb_get_local_majorant": (C.c_int, [VP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), c_float_p]),
    "dxb_get_dense_box": (C.c_int, [VP, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), c_float_p, c_float_p]),
    "dxb_device_majorant": (C.c_int, [VP, c_double_p, C.c_uint32, c_float_p]),
    "dxb_icrp_import": (C.c_int, [VP, c_u8_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_int, c_u8_p, c_u8_p, c_double_p, C.POINTER(VP)]),
    "dxb_icrp_plan": (C.c_int, [C.POINTER(VP), C.c_char_p, C.c_char_p, C.c_int, c_u8_p]),
    "dxb_icrp_luts": (C.c_int, [VP, c_u8_p, c_u8_p, c_double_p]),
    "dxb_icrp_destroy": (None, [VP]),
    "dxb_icrp_n_organs": (C.c_uint32, [VP]),
    "dxb_icrp_organ_name": (C.c_char_p, [VP, C.c_uint32]),
    "dxb_icrp_organ_density": (C.c_double, [VP, C.c_uint32]),
    "dxb_icrp_organ_medium": (C.c_uint32, [VP, C.c_uint32]),
    "dxb_icrp_n_media": (C.c_uint32, [VP]),
    "dxb_icrp_medium_name": (C.c_char_p, [VP, C.c_uint32]),
    "dxb_icrp_medium_composition": (C.c_int, [VP, C.c_uint32, c_u32_p, c_double_p, C.c_int]),
    "dxb_h5_create": (VP, []),
    "dxb_h5_open": (C.c_int, [C.POINTER(VP), C.c_char_p]),
    "dxb_h5_close": (None, [VP]),
    "dxb_h5_error": (C.c_char_p, [VP]),
    "dxb_h5_save": (C.c_int, [VP, C.c_char_p]),
    "dxb_h5_exists": (C.c_int, [VP, C.c_char_p]),
    "dxb_h5_make_group": (C.c_int, [VP, C.c_char_p]),
    "dxb_h5_list": (C.c_char_p, [VP, C.c_char_p]),
    "dxb_h5_put_dataset": (C.c_int, [VP, C.c_char_p, C.c_int, C.c_int, c_u64_p, VP, C.c_int]),
    "dxb_h5_put_strings": (C.c_int, [VP, C.c_char_p, C.c_uint64, C.POINTER(C.c_char_p)]),
    "dxb_h5_put_attribute": (C.c_int, [VP, C.c_char_p, C.c_char_p, C.c_int, C.c_int64, VP]),
    "dxb_h5_dataset_info": (C.c_int, [VP, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), c_u64_p, C.POINTER(C.c_int)]),
    "dxb_h5_dataset_read": (C.c_int, [VP, C.c_char_p, VP, C.c_uint64]),
    "dxb_h5_dataset_string": (C.c_char_p, [VP, C.c_char_p, C.c_uint64]),
    "dxb_h5_attribute_info": (C.c_int, [VP, C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "dxb_h5_attribute_read": (C.c_int, [VP, C.c_char_p, C.c_char_p, VP, C.c_uint64]),
    "dxb_h5_attribute_string": (C.c_char_p, [VP, C.c_char_p, C.c_char_p, C.c_uint64]),
    "dxb_load_scene": (C.c_int, [VP, C.c_char_p, c_u64_p, c_double_p, C.POINTER(C.c_uint32)]),
    "dxb_save_dose": (C.c_int, [VP, C.c_char_p, C.c_int, C.c_char_p]),
    "dxb_segment_ct": (C.c_int, [VP, c_double_p, C.c_uint64, C.POINTER(dxb_tube_desc), c_u8_p, c_double_p, C.POINTER(VP)]),
}

_lib = None


def load():
    """Load libdxmc_b200.so and bind every symbol of include/dxb.h.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make` (nvcc, sm_100a). libdxmc_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DxbError(RuntimeError):
    def __init__(self, code, where, detail=""):
        self.code = code
        name = STATUS_NAMES[code] if 0 <= code < len(STATUS_NAMES) else str(code)
        super().__init__(f"{where}: {name}" + (f" ({detail})" if detail else ""))

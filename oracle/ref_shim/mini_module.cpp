// Test infrastructure only: pybind11 entry point that registers the REFERENCE's own
// hot-path binding initialisers (compiled from /root/reference where they lie).
// No reference source is copied; this file only names the init functions declared
// in the reference's module.hpp (src/toast/_libtoast/module.hpp).
#include <module.hpp>
PYBIND11_MODULE(_toast_oracle, m) {
    register_aligned<toast::AlignedI8>(m, "AlignedI8");
    register_aligned<toast::AlignedU8>(m, "AlignedU8");
    register_aligned<toast::AlignedI64>(m, "AlignedI64");
    register_aligned<toast::AlignedF64>(m, "AlignedF64");
    init_intervals(m);
    init_template_offset(m);
    init_accelerator(m);
    init_ops_pointing_detector(m);
    init_ops_stokes_weights(m);
    init_ops_pixels_healpix(m);
    init_ops_mapmaker_utils(m);
    init_ops_noise_weight(m);
    init_ops_scan_map(m);
    init_map_cov(m);
    init_pixels(m);
}

// storage::b200 traits (include/gtb200/storage/b200.hpp) behind the reference's own storage::builder and data_store,
// and a registered spec through stencil::b200 -- PLAIN HOST CODE (g++, no nvcc): fill on the host, run on the device,
// read back; then the transfer rate of get_target_ptr() / host_view() for a 256x256x80 field.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <gridtools/stencil/cartesian.hpp>
#include <gridtools/storage/builder.hpp>
#include <gridtools/storage/sid.hpp>
#include <gtb200/stencil/b200.hpp>
#include <gtb200/storage/b200.hpp>

namespace gt = gridtools;
namespace st = gridtools::stencil;
using namespace gridtools::stencil;
using namespace gridtools::stencil::cartesian;

struct copy_functor {
    using in = in_accessor<0>;
    using out = inout_accessor<1>;
    using param_list = make_param_list<in, out>;
    template <class E>
    GT_FUNCTION static void apply(E eval) {
        eval(out()) = eval(in());
    }
};
GTB200_REGISTER_SPEC(gtb200::kernel::copy, copy_functor);

int main() {
    if (gtb_init(0) != GTB_OK) {
        std::printf("no device: %s\n", gtb_last_error());
        return 2;
    }
    int fails = 0;
    for (int ni : {37, 256}) {
        const int nj = ni == 37 ? 11 : 256, nk = ni == 37 ? 5 : 80;
        auto builder = gt::storage::builder<gt::storage::b200>.type<double>().dimensions(ni, nj, nk).halos(3, 3, 0);
        auto in = builder.initializer([](int i, int j, int k) { return i + 1000. * j + 1e6 * k; }).build();
        auto out = builder.value(-1).build();
        auto grid = st::make_grid(ni, nj, nk);
        st::run_single_stage(copy_functor(), st::b200<>(), grid, in, out);
        auto v = out->const_host_view(); // staged download, ordered behind the kernel on the legacy stream
        long bad = 0;
        for (int k = 0; k < nk; ++k)
            for (int j = 0; j < nj; ++j)
                for (int i = 0; i < ni; ++i)
                    bad += v(i, j, k) != i + 1000. * j + 1e6 * k;
        std::printf("copy %dx%dx%d through storage::b200 + stencil::b200: %s (%ld wrong)\n", ni, nj, nk, bad ? "FAILED" : "ok", bad);
        fails += bad != 0;
        if (ni == 256) { // transfer rates: modify on the host, fetch the target pointer (upload), touch the host view (download)
            const double mb = (double)in->info().length() * sizeof(double) / 1e6;
            for (int rep = 0; rep < 3; ++rep) {
                in->host_view()(0, 0, 0) = rep; // marks the device copy stale
                auto t0 = std::chrono::steady_clock::now();
                (void)in->get_const_target_ptr();
                gtb_stream_synchronize(nullptr);
                auto t1 = std::chrono::steady_clock::now();
                (void)out->get_target_ptr(); // marks the host copy stale
                auto t2 = std::chrono::steady_clock::now();
                (void)out->const_host_view();
                auto t3 = std::chrono::steady_clock::now();
                auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                std::printf("  %.1f MB: update_target %.2f ms (%.1f GB/s), update_host %.2f ms (%.1f GB/s)\n", mb, ms(t0, t1),
                    mb / ms(t0, t1), ms(t2, t3), mb / ms(t2, t3));
            }
        }
    }
    std::printf("%s\n", fails ? "SOME FAILED" : "ALL PASSED");
    return fails != 0;
}

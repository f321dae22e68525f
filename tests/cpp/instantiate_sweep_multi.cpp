// Compile-and-link check of the C++ wrappers of the pipelined sweep and the single-process multi-GPU group
// (bemb200::Sweep, bemb200::MultiGpu in include/bemb200.hpp).  The calls are instantiated, never executed (argc is never > 100):
// their executable twins are math_audio_b200/sweep.py (Sweep) and bem.py (MultiGpu), which the GPU tests drive.
#include "bemb200.hpp"
using namespace bemb200;
int main(int argc, char**) {
    if (argc > 100) {
        std::vector<Element> elements;
        std::vector<double> nodes;
        PhysicsParams physics(100.0, 343.0, 1.21, false);
        const GmresConfig config{10, 10, 1e-8, 0};
        Sweep sweep(0, elements, nodes);
        sweep.submit(physics, Complex64(0, 1), {}, config);
        sweep.set_block_jacobi(4);
        GmresSolution a = sweep.next();
        MultiGpu group({0, 1});
        MultiGpu::System system = group.build_tbem_system_with_beta(elements, nodes, physics, Complex64(0, 1));
        GmresSolution b = system.gmres(system.rhs, config);
        return static_cast<int>(a.iterations + b.iterations) + group.num_ranks() + static_cast<int>(sweep.num_dofs());
    }
    return 0;
}

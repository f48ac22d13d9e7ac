// ta_dense — the reference's dense matrix-multiply example (examples/gemm/ta_dense.cpp) written
// against include/tiledarray.hpp: same command line, same program text for the timed statement
// (c("m,n") = a("m,k") * b("k,n")), same report, evaluated by the B200 engine behind the C ABI.
//   ta_dense matrix_size block_size [repetitions]
// Build: g++ -std=c++17 -O2 -I include examples/ta_dense.cpp -L tiledarray_b200 -ltadev -o ta_dense
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <vector>

#include "tiledarray.hpp"

int main(int argc, char** argv) {
  int rc = 0;
  try {
    TiledArray::World& world = TiledArray::initialize(argc, argv);
    if (argc < 3) {
      std::cout << "Usage: ta_dense matrix_size block_size [repetitions]\n";
      return 0;
    }
    const long matrix_size = atol(argv[1]), block_size = atol(argv[2]);
    if (matrix_size <= 0 || block_size <= 0) { std::cerr << "Error: sizes must be greater than zero.\n"; return 1; }
    if ((matrix_size % block_size) != 0ul) { std::cerr << "Error: matrix size must be evenly divisible by block size.\n"; return 1; }
    const long repeat = (argc >= 4 ? atol(argv[3]) : 5);
    if (repeat <= 0) { std::cerr << "Error: number of repetitions must be greater than zero.\n"; return 1; }
    const std::size_t num_blocks = matrix_size / block_size, block_count = num_blocks * num_blocks;
    if (world.rank() == 0)
      std::cout << "TiledArray: dense matrix multiply test...\nEngine              = " << tadev_version()
                << "\nNumber of nodes     = " << world.size() << "\nMatrix size         = " << matrix_size << "x" << matrix_size
                << "\nBlock size          = " << block_size << "x" << block_size
                << "\nMemory per matrix   = " << double(matrix_size) * matrix_size * sizeof(double) / 1.0e9 << " GB\nNumber of blocks    = "
                << block_count << "\nAverage blocks/node = " << double(block_count) / double(world.size()) << "\n";
    std::vector<unsigned int> blocking;
    for (long i = 0l; i <= matrix_size; i += block_size) blocking.push_back(i);
    std::vector<TiledArray::TiledRange1> blocking2(2, TiledArray::TiledRange1(blocking.begin(), blocking.end()));
    TiledArray::TiledRange trange(blocking2.begin(), blocking2.end());
    const auto g = world.proc_grid(num_blocks, num_blocks, matrix_size, matrix_size);
    world.init_comm(g.proc_rows, g.proc_cols);

    const double gflops_per_call = 2.0 * double(matrix_size) * matrix_size * matrix_size / 1.0e9;
    {
      TiledArray::TArrayD a(world, trange), b(world, trange), c(world, trange);
      a.fill(1.0);
      b.fill(1.0);
      world.sync();
      if (world.rank() == 0) std::cout << "Starting iterations: \n";
      std::vector<double> durations;
      for (int i = 0; i < repeat; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        c("m,n") = a("m,k") * b("k,n");
        world.sync();
        const double time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        durations.push_back(time);
        if (world.rank() == 0) std::cout << "Iteration " << i + 1 << "   time=" << time << "   GFLOPS=" << gflops_per_call / time << "\n";
      }
      if (world.rank() == 0) {
        double mean = 0, mean_rec = 0;
        for (double d : durations) { mean += d; mean_rec += 1.0 / d; }
        mean /= durations.size(); mean_rec /= durations.size();
        std::sort(durations.begin(), durations.end());
        const double median = durations[durations.size() / 2];
        std::cout << "Average wall time   = " << mean << " s\nAverage GFLOPS      = " << gflops_per_call * mean_rec
                  << "\nMedian wall time   = " << median << " s\nMedian GFLOPS      = " << gflops_per_call / median << "\n";
        // the reference's device example verifies the result (examples/device/ta_dense_device.cpp): every element == N
        const auto t = c.find(0).get();
        bool ok = true;
        for (size_t i = 0; i < t.size(); ++i) ok = ok && t[i] == double(matrix_size);
        std::cout << "Verification        = " << (ok ? "passed" : "FAILED") << "\n";
        if (!ok) rc = 1;
      }
    }
    TiledArray::finalize();
  } catch (TiledArray::Exception& e) {
    std::cerr << "!! TiledArray exception: " << e.what() << "\n";
    rc = 1;
  } catch (std::exception& e) {
    std::cerr << "!! std exception: " << e.what() << "\n";
    rc = 1;
  }
  return rc;
}

// tmpfs_write_bench.cpp -- how fast can T threads fill a NEW 648 MB file on /dev/shm (the output of the FASTQ pipeline)?
//   mode 0: ftruncate + mmap + first-touch stores     mode 1: fallocate first, then the same     mode 2: pwrite per thread
// g++ -O2 -pthread -o tmpfs_write_bench tmpfs_write_bench.cpp;  ./tmpfs_write_bench <threads> <mode>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
    const size_t bytes = 648ull << 20;
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const int mode = argc > 2 ? atoi(argv[2]) : 0;  // 0 plain, 1 fallocate first, 2 pwrite per thread 1MB blocks
    const char *path = "/dev/shm/kbbq_tmpfs_write_bench.bin";
    for (int rep = 0; rep < 3; ++rep) {
        unlink(path);
        int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
        double t0 = now();
        double tf = 0;
        if (mode == 1) { if (fallocate(fd, 0, 0, bytes)) perror("fallocate"); tf = now() - t0; }
        else if (mode != 2) { if (ftruncate(fd, bytes)) perror("ftruncate"); }
        if (mode == 2) {
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back([=] {
                std::vector<char> buf(1 << 20, 'x');
                size_t lo = bytes * t / T, hi = bytes * (t + 1) / T;
                for (size_t o = lo; o < hi; o += buf.size()) { size_t n = std::min(buf.size(), hi - o); if (pwrite(fd, buf.data(), n, o) < 0) perror("pwrite"); }
            });
            for (auto &x : th) x.join();
        } else {
            char *m = (char *)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            std::vector<std::thread> th;
            for (int t = 0; t < T; ++t) th.emplace_back([=] {
                size_t lo = bytes * t / T, hi = bytes * (t + 1) / T;
                memset(m + lo, 'x', hi - lo);
            });
            for (auto &x : th) x.join();
            munmap(m, bytes);
        }
        double dt = now() - t0;
        printf("mode %d threads %d: %.1f ms (fallocate %.1f ms) %.1f GB/s\n", mode, T, dt * 1e3, tf * 1e3, bytes / dt / 1e9);
        close(fd);
    }
    unlink(path);
}

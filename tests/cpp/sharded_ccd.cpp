// Multi-GPU ccd() from plain C++ through the C ABI (include/sccd.h), one process per GPU.
//
//   <launcher> ./sharded_ccd mesh.bin rendezvous_dir
//
// Launch it the way torchrun / mpirun launch anything: the launcher sets RANK, WORLD_SIZE and
// LOCAL_RANK (OMPI_COMM_WORLD_* are read as well), e.g.
//   python -m torch.distributed.run --no-python --nproc-per-node 8 ./sharded_ccd mesh.bin /tmp/rdv
// Rank 0 makes the NCCL id (sccd_comm_get_unique_id) and publishes it as a file in the
// rendezvous directory; the others wait for the file -- any other broadcast (MPI_Bcast, a TCP
// store) does the same job.  Every rank uploads the same mesh and calls sccd_ccd_sharded.
//
// mesh.bin: int64 nV, nE, nF, then V0, V1 (nV x 3 float64, column-major), E (nE x 2 int32,
// column-major), F (nF x 3 int32, column-major) -- the layout the reference's ccd() takes.
// Prints one line per rank:  rank R/W toi=<%.17g> pairs=<vf> <ee> ms=<device time of the step>
#include "sccd.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static int env_int(const char* a, const char* b, int dflt)
{
    const char* v = getenv(a);
    if (!v && b)
        v = getenv(b);
    return v ? atoi(v) : dflt;
}

#define CHECK(call)                                                                      \
    do {                                                                                 \
        const int rc_ = (call);                                                          \
        if (rc_ != SCCD_OK) {                                                            \
            fprintf(stderr, "rank %d: %s failed (%d): %s\n", rank, #call, rc_,           \
                    ctx ? sccd_last_error(ctx) : "");                                    \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char** argv)
{
    const int rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", 0);
    const int world = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1);
    const int local = env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", rank);
    sccd_ctx* ctx = nullptr;
    if (argc < 3) {
        fprintf(stderr, "usage: sharded_ccd mesh.bin rendezvous_dir [steps]\n");
        return 2;
    }
    const int steps = argc > 3 ? atoi(argv[3]) : 1;
    // ---- mesh
    FILE* f = fopen(argv[1], "rb");
    if (!f) {
        fprintf(stderr, "cannot open %s\n", argv[1]);
        return 2;
    }
    int64_t n[3];
    if (fread(n, 8, 3, f) != 3)
        return 2;
    std::vector<double> V0(3 * n[0]), V1(3 * n[0]);
    std::vector<int32_t> E(2 * n[1]), F(3 * n[2]);
    if (fread(V0.data(), 8, V0.size(), f) != V0.size() || fread(V1.data(), 8, V1.size(), f) != V1.size()
        || fread(E.data(), 4, E.size(), f) != E.size() || fread(F.data(), 4, F.size(), f) != F.size())
        return 2;
    fclose(f);

    CHECK(sccd_create(local, nullptr, &ctx));
    // ---- rendezvous: rank 0's NCCL id travels through a file
    unsigned char id[SCCD_UNIQUE_ID_BYTES];
    const std::string path = std::string(argv[2]) + "/sccd_nccl_id";
    if (world > 1) {
        if (rank == 0) {
            CHECK(sccd_comm_get_unique_id(id));
            const std::string tmp = path + ".tmp";
            FILE* o = fopen(tmp.c_str(), "wb");
            if (!o || fwrite(id, 1, sizeof(id), o) != sizeof(id))
                return 2;
            fclose(o);
            rename(tmp.c_str(), path.c_str()); // atomic: readers never see a partial id
        } else {
            for (int tries = 0;; tries++) {
                FILE* i = fopen(path.c_str(), "rb");
                if (i) {
                    const size_t got = fread(id, 1, sizeof(id), i);
                    fclose(i);
                    if (got == sizeof(id))
                        break;
                }
                if (tries > 6000) {
                    fprintf(stderr, "rank %d: no NCCL id after 60 s\n", rank);
                    return 3;
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(10));
            }
        }
    }
    CHECK(sccd_comm_create(ctx, world > 1 ? id : nullptr, rank, world));
    CHECK(sccd_upload_mesh(ctx, V0.data(), V1.data(), n[0], E.data(), n[1], F.data(), n[2], 0));
    double toi = 1.0;
    sccd_stats st;
    for (int s = 0; s < steps; s++)
        CHECK(sccd_ccd_sharded(ctx, 0.0, -1, 1e-6, 1, &toi)); // tests/test_narrow_phase.cu:41-45
    CHECK(sccd_get_stats(ctx, &st));
    printf("rank %d/%d toi=%.17g pairs=%lld %lld ms=%.3f\n", rank, world, toi,
           (long long)st.n_pairs[0], (long long)st.n_pairs[1], st.ms_total);
    CHECK(sccd_comm_destroy(ctx));
    sccd_destroy(ctx);
    return 0;
}

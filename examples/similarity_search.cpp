// similarity_search.cpp -- the reference's examples/similarity_search.rs on the B200 engine, written
// against the C ABI (include/mvf_b200.h) only -- exactly what a foreign-language binding sees.
//
//   g++ -std=c++17 -Iinclude examples/similarity_search.cpp -Lmetrovector_b200 -lmvf_b200 \
//       -Wl,-rpath,$PWD/metrovector_b200 -o similarity_search && ./similarity_search [n_gpus]
//
// With an argument > 1 the same program drives that many GPUs from this one process (mvf_gpu_init_multi: the
// space is row-sharded over the devices, every search uses all of them); nothing else in the program changes.
//
// Same dataset (three clusters of 20 vectors, examples/similarity_search.rs:43-76), same queries
// (:103-108), k = 5 (:116).  Prints the k NEAREST vectors per query (the documented intent; the
// reference's loop as written keeps the k farthest).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "mvf_b200.h"

static void check(int rc, const char* what) {
    if (rc != MVF_OK) { std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, mvf_last_error()); std::exit(1); }
}

int main(int argc, char** argv) {
    const int n_gpus = argc > 1 ? std::atoi(argv[1]) : 1;
    std::printf("=== MVF Similarity Search Example (B200) ===\n\n");
    std::vector<float> vectors;
    const float centers[3][4] = {{1, 1, 1, 1}, {5, 5, 5, 5}, {-2, 3, 0, 4}};
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < 20; i++) {
            const float noise = (float)i * 0.1f;
            if (c < 2) { vectors.insert(vectors.end(), {centers[c][0] + noise, centers[c][1] - noise, centers[c][2] + noise * 0.5f, centers[c][3] - noise * 0.5f}); }
            else { vectors.insert(vectors.end(), {-2.0f + noise, 3.0f - noise, 0.0f + noise, 4.0f - noise * 0.5f}); }
        }
    std::printf("Created dataset with %zu vectors in 3 clusters\n", vectors.size() / 4);

    // build + save (MvfBuilder::add_vector_space / add_vectors / build / save, src/builder.rs)
    mvf_builder* builder = nullptr;
    check(mvf_builder_new(&builder), "mvf_builder_new");
    check(mvf_builder_add_vector_space(builder, "clustered_data", 4, MVF_DENSE, MVF_L2, MVF_FLOAT32, nullptr), "add_vector_space");
    check(mvf_builder_add_vectors_f32(builder, "clustered_data", vectors.data(), vectors.size() / 4, 4), "add_vectors");
    const std::string path = "/tmp/similarity_search.mvf";
    check(mvf_builder_save(builder, path.c_str()), "save");
    mvf_builder_free(builder);

    // open (MvfReader::open), look the space up, upload it once
    mvf_file* file = nullptr;
    check(mvf_file_open(path.c_str(), &file), "mvf_file_open");
    const char* name = nullptr;
    check(mvf_file_vector_space_name(file, 0, &name), "vector_space_names");
    mvf_space* space = nullptr;
    check(mvf_file_vector_space(file, name, &space), "vector_space");
    mvf_gpu* gpu = nullptr;
    if (n_gpus > 1) {
        std::vector<int> devices(n_gpus);
        for (int i = 0; i < n_gpus; i++) devices[i] = i;
        check(mvf_gpu_init_multi(devices.data(), n_gpus, &gpu), "mvf_gpu_init_multi");
    } else {
        check(mvf_gpu_init(0, &gpu), "mvf_gpu_init");
    }
    mvf_gpu_space* resident = nullptr;
    check(mvf_gpu_space_upload(gpu, space, &resident), "mvf_gpu_space_upload");

    const float queries[4][4] = {{1, 1, 1, 1}, {5, 5, 5, 5}, {-2, 3, 0, 4}, {0, 0, 0, 0}};
    const char* labels[4] = {"Near cluster 1", "Near cluster 2", "Near cluster 3", "At origin"};
    const uint32_t k = 5;
    for (int qi = 0; qi < 4; qi++) {
        std::printf("\n=== Query: %s ===\nQuery vector: [%g, %g, %g, %g]\nTop %u most similar vectors:\n", labels[qi], queries[qi][0],
                    queries[qi][1], queries[qi][2], queries[qi][3], k);
        uint64_t ids[k]; float scores[k];
        check(mvf_gpu_search(resident, queries[qi], MVF_FLOAT32, 1, k, MVF_METRIC_SPACE_DEFAULT, ids, scores, nullptr), "mvf_gpu_search");
        int counts[3] = {0, 0, 0};
        for (uint32_t r = 0; r < k; r++) {
            float row[4];
            check(mvf_space_get_vector_f32(space, ids[r], row, 4), "get_vector");
            std::printf("  %u. Vector %llu (distance: %.3f): [%g, %g, %g, %g]\n", r + 1, (unsigned long long)ids[r], scores[r], row[0],
                        row[1], row[2], row[3]);
            counts[row[0] < 2.0f ? 0 : row[0] > 4.0f ? 1 : 2]++;      // analyze_clusters, :178-198
        }
        std::printf("  Cluster distribution: C1=%d, C2=%d, C3=%d\n", counts[0], counts[1], counts[2]);
    }
    mvf_gpu_space_free(resident);
    mvf_gpu_shutdown(gpu);
    mvf_space_release(space);
    mvf_file_close(file);
    std::remove(path.c_str());
    return 0;
}

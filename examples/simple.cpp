// simple.cpp -- the reference's examples/simple.rs on the B200 engine, written against the C ABI
// (include/mvf_b200.h) only.
//
//   g++ -std=c++17 -Iinclude examples/simple.cpp -Lmetrovector_b200 -lmvf_b200 -Wl,-rpath,$PWD/metrovector_b200 -o simple
//   ./simple
//
// Same five vectors (examples/simple.rs:15-21), same query (:70), same steps: build a file, read it back, print
// the vectors, then the nearest one.  Steps 1-3 are host work (builder + reader); step 4 is the 1-NN loop of
// examples/simple.rs:73-94 as ONE search with k = total_vectors (all distances, which the reference prints) whose
// first result is the best match -- strict '<' in the reference means the lowest index wins a tie, and so does
// the engine.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "mvf_b200.h"

static void check(int rc, const char* what) {
    if (rc != MVF_OK) { std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, mvf_last_error()); std::exit(1); }
}
static const char* data_type_name(int dt) {
    switch (dt) { case MVF_FLOAT32: return "Float32"; case MVF_FLOAT16: return "Float16"; case MVF_INT8: return "Int8"; case MVF_UINT8: return "UInt8"; default: return "?"; }
}

int main() {
    std::printf("=== MVF Simple Example (B200) ===\n\n");
    const std::vector<float> vectors = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 2, 4, 6, 8, 1, 3, 5, 7};
    const size_t n = vectors.size() / 4;
    std::printf("1. Building MVF file with %zu vectors\n", n);
    mvf_builder* builder = nullptr;
    check(mvf_builder_new(&builder), "mvf_builder_new");
    check(mvf_builder_add_vector_space(builder, "embeddings", 4, MVF_DENSE, MVF_L2, MVF_FLOAT32, nullptr), "add_vector_space");
    check(mvf_builder_add_vectors_f32(builder, "embeddings", vectors.data(), n, 4), "add_vectors");
    const std::string path = "/tmp/simple_example.mvf";
    check(mvf_builder_save(builder, path.c_str()), "save");
    mvf_builder_free(builder);
    std::printf("   Written to: \"%s\"\n", path.c_str());

    std::printf("\n2. Reading MVF file\n");
    mvf_file* file = nullptr;
    check(mvf_file_open(path.c_str(), &file), "mvf_file_open");
    uint16_t version = 0; size_t spaces = 0;
    check(mvf_file_version(file, &version), "version");
    check(mvf_file_num_vector_spaces(file, &spaces), "num_vector_spaces");
    std::printf("   File version: %u\n   Vector spaces: %zu\n", (unsigned)version, spaces);
    const char* first = nullptr;
    check(mvf_file_vector_space_name(file, 0, &first), "vector_space_names");
    mvf_space* space = nullptr;
    check(mvf_file_vector_space(file, first, &space), "vector_space");
    const char* name = nullptr; uint32_t dim = 0; uint64_t total = 0; int dtype = 0;
    check(mvf_space_name(space, &name), "name");
    check(mvf_space_dimension(space, &dim), "dimension");
    check(mvf_space_total_vectors(space, &total), "total_vectors");
    check(mvf_space_data_type(space, &dtype), "data_type");
    std::printf("   Space name: '%s'\n   Dimensions: %u\n   Total vectors: %llu\n   Data type: %s\n", name, dim,
                (unsigned long long)total, data_type_name(dtype));

    std::printf("\n3. Reading vectors:\n");
    for (uint64_t i = 0; i < total; i++) {
        float row[4];
        check(mvf_space_get_vector_f32(space, i, row, 4), "get_vector");
        std::printf("   Vector %llu: [%.1f, %.1f, %.1f, %.1f]\n", (unsigned long long)i, row[0], row[1], row[2], row[3]);
    }

    std::printf("\n4. Similarity search:\n");
    const float query[4] = {2.5f, 4.5f, 6.5f, 8.5f};
    std::printf("   Query vector: [%.1f, %.1f, %.1f, %.1f]\n", query[0], query[1], query[2], query[3]);
    mvf_gpu* gpu = nullptr;
    check(mvf_gpu_init(0, &gpu), "mvf_gpu_init");
    mvf_gpu_space* resident = nullptr;
    check(mvf_gpu_space_upload(gpu, space, &resident), "mvf_gpu_space_upload");
    std::vector<uint64_t> ids(total);
    std::vector<float> scores(total);
    check(mvf_gpu_search(resident, query, MVF_FLOAT32, 1, (uint32_t)total, MVF_METRIC_SPACE_DEFAULT, ids.data(), scores.data(), nullptr),
          "mvf_gpu_search");
    std::vector<float> by_index(total, 0.0f);
    for (uint64_t r = 0; r < total; r++) by_index[ids[r]] = scores[r];
    for (uint64_t i = 0; i < total; i++) std::printf("   Vector %llu distance: %.2f\n", (unsigned long long)i, by_index[i]);
    float best[4];
    check(mvf_space_get_vector_f32(space, ids[0], best, 4), "get_vector");
    std::printf("   Best match: Vector %llu (distance: %.2f)\n   Result: [%.1f, %.1f, %.1f, %.1f]\n", (unsigned long long)ids[0], scores[0],
                best[0], best[1], best[2], best[3]);

    mvf_gpu_space_free(resident);
    mvf_gpu_shutdown(gpu);
    mvf_space_release(space);
    mvf_file_close(file);
    std::remove(path.c_str());
    return 0;
}

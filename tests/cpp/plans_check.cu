// Host-side checks of the layout / planning helpers of the fused mean-field path (no GPU needed: nothing here launches).
//   blur_multi_plan  every lattice blurs each of its d+1 axes exactly once, in order, in at most RSS_BLUR_FUSE(3)-axis chunks,
//                    in consecutive phases starting at phase 0 (the ping / pong parity of the result depends on it)
//   q_row            the rotation of the Q-tile rows is a permutation inside every warp's 32 rows
//   pt_index         the point-major (corner, point) index is a bijection onto [0, D1 * TILE_POINTS)
//   fused_tile_map   tile counts of image and 1-D point sets
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "meanfield.cuh"

using namespace rss;

#define CHECK(c)                                                        \
    do {                                                                \
        if (!(c)) { printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } \
    } while (0)

static int check_plan(int K, const int* d1, const uint32_t* vcap, int G) {
    BlurMultiArgs a{};
    a.K = K;
    for (int k = 0; k < K; k++) { a.d1[k] = d1[k]; a.vcap[k] = vcap[k]; }
    int phases_of[FUSED_MAX_LAT] = {0};
    const int phases = blur_multi_plan(a, G, phases_of);
    CHECK(phases >= 1 && phases <= BLUR_MAX_PHASES && phases == a.phases);
    int most = 0;
    for (int k = 0; k < K; k++) {
        int sum = 0, np = 0;
        bool ended = false;
        for (int p = 0; p < BLUR_MAX_PHASES; p++) {
            const int f = a.fuse[k][p];
            CHECK(f >= 0 && f <= 3);
            if (f == 0) { ended = true; continue; }
            CHECK(!ended);      // consecutive phases from phase 0 on
            CHECK(p < phases);
            sum += f;
            np++;
        }
        CHECK(sum == d1[k]);    // every axis exactly once
        CHECK(np == phases_of[k]);
        most = np > most ? np : most;
    }
    CHECK(most == phases);      // no empty phase
    for (int k = K; k < FUSED_MAX_LAT; k++)
        for (int p = 0; p < BLUR_MAX_PHASES; p++) CHECK(a.fuse[k][p] == 0);
    return 0;
}

int main() {
    // the keyframe CRF (Gaussian 3-D + bilateral 5-D, 5 channel groups) and a range of other shapes / table sizes
    const uint32_t caps[] = {1u << 10, 1u << 16, 1u << 18, 1u << 22, 1u << 24};
    for (int G = 1; G <= 6; G++)
        for (int da = 2; da <= 8; da++)
            for (uint32_t ca : caps) {
                const int d1[2] = {da, 0};
                const uint32_t vc[2] = {ca, 0};
                if (check_plan(1, d1, vc, G)) return 1;
                for (int db = 2; db <= 8; db++)
                    for (uint32_t cb : caps) {
                        const int e1[2] = {da, db};
                        const uint32_t wc[2] = {ca, cb};
                        if (check_plan(2, e1, wc, G)) return 1;
                    }
            }
    {   // the benchmark's signature: 4 axes spread over the bilateral lattice's 3 phases
        BlurMultiArgs a{};
        a.K = 2; a.d1[0] = 4; a.d1[1] = 6; a.vcap[0] = 65536; a.vcap[1] = 16384;
        int po[FUSED_MAX_LAT];
        CHECK(blur_multi_plan(a, 5, po) == 3);
        CHECK(a.fuse[0][0] == 2 && a.fuse[0][1] == 1 && a.fuse[0][2] == 1);
        CHECK(a.fuse[1][0] == 2 && a.fuse[1][1] == 2 && a.fuse[1][2] == 2);
    }
    // q_row: a permutation of every warp's 32 rows
    for (int w = 0; w < TILE_POINTS / 32; w++) {
        unsigned seen = 0;
        for (int l = 0; l < 32; l++) {
            const int q = q_row(32 * w + l);
            CHECK(q >= 32 * w && q < 32 * w + 32);
            seen |= 1u << (q & 31);
        }
        CHECK(seen == 0xffffffffu);
    }
    // pt_index: bijection
    for (int D1 = 1; D1 <= 8; D1++) {
        std::vector<int> hit(D1 * TILE_POINTS, 0);
        for (int j = 0; j < D1; j++)
            for (int lp = 0; lp < TILE_POINTS; lp++) {
                const int i = pt_index(D1, j, lp);
                CHECK(i >= 0 && i < D1 * TILE_POINTS);
                hit[i]++;
            }
        for (int v : hit) CHECK(v == 1);
    }
    // tile maps
    {
        const TileMap m = fused_tile_map(640 * 480, 640, 480);
        CHECK(m.W == 640 && m.TW == 32 && m.TH == 8 && m.tiles_x == 20 && m.ntiles == 1200);
        const TileMap n = fused_tile_map(1000, 0, 0);
        CHECK(n.W == 0 && n.ntiles == 4 && n.perm == nullptr);
        const TileMap o = fused_tile_map(100 * 50, 100, 50);
        CHECK(o.tiles_x == 4 && o.ntiles == 4 * 7);
    }
    CHECK(fused_signature_supported(5, 4, 6) && fused_signature_supported(3, 7, 0) && !fused_signature_supported(7, 4, 6));
    printf("plans_check OK\n");
    return 0;
}

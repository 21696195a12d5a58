// host_selftest.cpp -- drives the library-level API of the C++ host mirror and prints JSON that
// tests/test_host_mirror.py compares with the oracle and the reference's goldens.
//   host_selftest <file.tsp>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "teeline_host.hpp"

using namespace teeline;
using namespace teeline::tsp;

static void print_ids(const char *key, const std::vector<size_t> &v, bool comma = true)
{
    std::printf("\"%s\":[", key);
    for (size_t k = 0; k < v.size(); ++k) std::printf("%s%zu", k ? "," : "", v[k]);
    std::printf("]%s", comma ? "," : "");
}

static uint32_t bits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    try {
        auto data = tsplib::read_from_file(argv[1]);
        if (data.is_err()) { std::fprintf(stderr, "%s\n", data.error.c_str()); return 1; }
        tsplib::TspLibData d = data.unwrap();
        auto dmr = d.distance_matrix();
        if (dmr.is_err()) { std::fprintf(stderr, "%s\n", dmr.error.c_str()); return 1; }
        TspProblem problem{d.cities, dmr.unwrap()};
        const DistanceMatrix &dm = problem.distances;
        const size_t n = d.cities.size();
        std::printf("{\"name\":\"%s\",\"n\":%zu,\"explicit\":%s,", d.name.c_str(), n, d.has_explicit_weights() ? "true" : "false");

        // DistanceMatrix::distances(): FNV-1a over the bit patterns + a few entries
        uint64_t h = 1469598103934665603ull;
        for (float f : dm.distances()) { h ^= bits(f); h *= 1099511628211ull; }
        std::printf("\"matrix_len\":%zu,\"matrix_fnv\":\"%016llx\",", dm.len(), (unsigned long long)h);
        std::printf("\"d_first_last_bits\":[%u,%u],", bits(*dm.distance_between(d.cities[0].id, d.cities[n - 1].id)),
                    bits(*dm.distance_by_pos(n - 1, 0)));
        std::printf("\"d_unknown_is_none\":%s,", dm.distance_between(d.cities[0].id, 987654321).has_value() ? "false" : "true");

        // nearest() for the first and last city, k = 3 and 5; build_candidates k = 5
        for (size_t k : {3, 5}) {
            for (size_t which : {(size_t)0, n - 1}) {
                NearestResult r = dm.nearest(d.cities[which], k);
                std::vector<size_t> ids;
                for (auto &it : r.nearest()) ids.push_back(it.point.id);
                char key[64];
                std::snprintf(key, sizeof key, "nearest_k%zu_%s", k, which == 0 ? "first" : "last");
                print_ids(key, ids);
            }
        }
        KDPoint stranger = KDPoint::new_with_id(987654321, 0.f, 0.f);
        std::printf("\"nearest_unknown_len\":%zu,", dm.nearest(stranger, 3).nearest().size());
        auto cand = build_candidates(d.cities, dm, 5);
        print_ids("candidates_first", cand[d.cities[0].id]);
        print_ids("candidates_last", cand[d.cities[n - 1].id]);

        // tour lengths: identity order, with an unknown id, batch of two
        std::vector<size_t> ident;
        for (auto &c : d.cities) ident.push_back(c.id);
        std::printf("\"len_identity_bits\":%u,", bits(dm.tour_length(ident)));
        std::vector<size_t> bad = ident;
        bad[n / 2] = 987654321;
        std::printf("\"len_unknown_id\":%.1f,", dm.tour_length(bad));
        std::vector<size_t> two = ident;
        std::vector<size_t> rev(ident.rbegin(), ident.rend());
        two.insert(two.end(), rev.begin(), rev.end());
        auto lens = dm.tour_lengths(two, 2);
        std::printf("\"len_batch_bits\":[%u,%u],", bits(lens[0]), bits(lens[1]));

        // solvers through solve_with_context, with a progress sink on 2-opt and or-opt
        AppOptions opts;
        auto nn = solve_problem(Solvers::NearestNeighbor, problem, opts);
        std::printf("\"nn_total\":\"%.5f\",", nn.unwrap().total);
        print_ids("nn_route", nn.unwrap().route());
        size_t updates = 0, dones = 0;
        float last_total = 0.f;
        std::vector<size_t> last_route;
        ProgressSender sink = [&](const ProgressMessage &m) {
            if (m.kind == ProgressMessage::PathUpdate) { ++updates; last_total = m.total; last_route = m.route; }
            if (m.kind == ProgressMessage::Done) ++dones;
        };
        auto seed = nn.unwrap().route();
        auto two_opt_sol = solve_with_context(Solvers::TwoOpt, problem, opts, &sink, &seed);
        std::printf("\"two_opt_total\":\"%.5f\",\"two_opt_updates\":%zu,\"two_opt_dones\":%zu,\"two_opt_last_update_is_final\":%s,",
                    two_opt_sol.unwrap().total, updates, dones, last_route == two_opt_sol.unwrap().route() ? "true" : "false");
        print_ids("two_opt_route", two_opt_sol.unwrap().route());
        updates = dones = 0;
        auto or_sol = solve_with_context(Solvers::OrOpt, problem, opts, &sink, &seed);
        std::printf("\"or_opt_total\":\"%.5f\",\"or_opt_updates\":%zu,\"or_opt_last_total_bits\":%u,\"or_opt_total_bits\":%u,",
                    or_sol.unwrap().total, updates, bits(last_total), bits(or_sol.unwrap().total));
        print_ids("or_opt_route", or_sol.unwrap().route());
        auto noseed = solve_with_context(Solvers::TwoOpt, problem, opts, nullptr, nullptr);
        std::printf("\"two_opt_noseed_total\":\"%.5f\",", noseed.unwrap().total);
        auto best = solve_with_context(Solvers::TwoOptBest, problem, opts, nullptr, &seed);
        std::printf("\"two_opt_best_total\":\"%.5f\",", best.unwrap().total);
        print_ids("two_opt_best_route", best.unwrap().route());
        auto sa = solve_problem(Solvers::SimulatedAnnealing, problem, opts);
        std::printf("\"sa_is_err\":%s,", sa.is_err() ? "true" : "false");
        AppOptions badopts;
        badopts.heuristic = HeuristicOptions{};
        badopts.heuristic->n_nearest = 0;
        auto bad_nn = solve_problem(Solvers::NearestNeighbor, problem, badopts);
        std::printf("\"n_nearest_zero_error\":\"%s\",", bad_nn.error.c_str());
        std::vector<size_t> dup = seed;
        dup[1] = dup[0];
        std::printf("\"validate_dup\":\"%s\"}\n", validate_tour(dup, d.cities).error.c_str());
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "panicked: %s\n", e.what());
        return 101;
    }
}

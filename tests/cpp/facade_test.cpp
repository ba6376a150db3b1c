// Drop-in check of include/PiXiuCtrl.hpp: the reference's own t_PiXiuCtrl scenario
// (/root/reference/src/proj/PiXiuCtrl.cpp:116-255 — max-length records, CRUD differential against
// std::map over {A..E}, prefix iteration), written against the reference's call shapes, scaled down
// because every single-record call is a GPU batch of one.  Exit code 0 = pass.
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "PiXiuCtrl.hpp"

#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) {                                                   \
            fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                 \
        }                                                             \
    } while (0)

static std::string drain(PXSGen *gen) {
    std::string out;
    uint8_t rv;
    while (gen->operator()(rv)) out.push_back((char) rv);
    PXSGen_free(gen);
    return out;
}

int main() {
    srand(19950207);  // PiXiuCtrl.cpp:119
    PiXiuCtrl ctrl;
    ctrl.init_prop();
    CHECK(ctrl.store != nullptr);

    // max-length key-only record: 65,533 raw bytes -> 65,535 decoded (PiXiuCtrl.cpp:121-153)
    {
        std::vector<uint8_t> k(65533, 'K');
        CHECK(ctrl.setitem(k.data(), (int) k.size(), NULL, 0) == 0);
        CHECK(ctrl.contains(k.data(), (int) k.size()));
        std::string d = drain(ctrl.getitem(k.data(), (int) k.size()));
        CHECK(d.size() == 65535 && (uint8_t) d[65533] == 251 && d[65534] == 0);
        k.push_back('K');
        CHECK(ctrl.setitem(k.data(), (int) k.size(), NULL, 0) == PIXIU_ETOOLONG);  // reference: assert
        k.pop_back();
        CHECK(ctrl.delitem(k.data(), (int) k.size()) == 0);
    }
    // CRUD differential vs std::map (PiXiuCtrl.cpp:176-226)
    std::map<std::string, std::string> model;
    auto rnd = [](int lo, int hi) {
        int n = lo + rand() % (hi - lo + 1);
        std::string s;
        for (int i = 0; i < n; i++) s.push_back((char) ('A' + rand() % 5));
        return s;
    };
    for (int it = 0; it < 1500; it++) {
        std::string k = rnd(1, 5), v = rnd(1, 50);
        int op = rand() % 4;
        if (op <= 1) {
            int rc = ctrl.setitem((uint8_t *) k.data(), (int) k.size(), (uint8_t *) v.data(), (int) v.size());
            CHECK(rc == (model.count(k) ? CBT_SET_REPLACE : 0));
            model[k] = v;
        } else if (op == 2) {
            int rc = ctrl.delitem((uint8_t *) k.data(), (int) k.size());
            CHECK(rc == (model.count(k) ? 0 : CBT_DEL_NOT_FOUND));
            model.erase(k);
        } else {
            CHECK(ctrl.contains((uint8_t *) k.data(), (int) k.size()) == (model.count(k) > 0));
            PXSGen *g = ctrl.getitem((uint8_t *) k.data(), (int) k.size());
            if (model.count(k)) {
                CHECK(g != NULL);
                std::string want = k + "\xfb" + std::string(1, '\0') + model[k] + "\xfb\x02";
                CHECK(drain(g) == want);
            } else {
                CHECK(g == NULL);
            }
        }
    }
    // prefix iteration (PiXiuCtrl.cpp:228-255): the order is byte-lexicographic on esc(key) 251 0, i.e. a key
    // sorts AFTER the keys it is a prefix of (251 > 'A'); the reference's own test sorts before comparing
    for (std::string prefix : {std::string(""), std::string("A"), std::string("CD"), std::string("EEEEEE")}) {
        std::vector<std::string> want;
        for (auto &kv : model)
            if (kv.first.compare(0, prefix.size(), prefix) == 0)
                want.push_back(kv.first + "\xfb" + std::string(1, '\0') + kv.second + "\xfb\x02");
        std::sort(want.begin(), want.end());  // std::string compares as unsigned bytes
        std::vector<std::string> got;
        CBTGen *it = ctrl.iter((uint8_t *) prefix.data(), (int) prefix.size());
        if (it) {
            PXSGen *g;
            while (it->operator()(g)) got.push_back(drain(g));
            CBTGen_free(it);
        }
        CHECK(got == want);
    }
    // consume_repr keeps visible bytes only (PiXiuStr.h:200-211)
    {
        const char *k = "repr", *v = "a b\tc";
        CHECK(ctrl.setitem((uint8_t *) k, 4, (uint8_t *) v, 5) == 0);
        char *r = ctrl.getitem((uint8_t *) k, 4)->consume_repr();
        CHECK(std::string(r) == "reprabc");
        free(r);
    }
    ctrl.free_prop();
    printf("facade_test ok (%zu live keys)\n", model.size());
    return 0;
}

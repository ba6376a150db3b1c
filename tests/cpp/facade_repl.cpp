// The reference's REPL (/root/reference/src/main.cpp:40-76, the Release branch of its main) written against
// include/PiXiuCtrl.hpp: `SET k::v` stores (the value keeps the "::" like in the reference), `GET k` prints the visible
// bytes of the record, every SET prints the bytes it saved (main.cpp:67-70) and the running total, `~` quits.
// tests/test_repl_transcript.py feeds the same command file to this program and to the reference's own binary
// (oracle/_ref/pixiu_repl) and compares the two transcripts byte for byte.
#include <cstdio>
#include <iostream>
#include <string>

#include "PiXiuCtrl.hpp"

int main() {
    PiXiuCtrl ctrl;
    ctrl.config.strict251 = 1;   // the reference's bytes, so that the saved-bytes line matches in every case
    ctrl.init_prop();
    if (!ctrl.store) return 2;

    std::string cmd;
    long total_diff = 0;
    while (true) {
        std::cout << "Command: ";
        if (!std::getline(std::cin, cmd)) break;
        if (!cmd.empty() && cmd[0] == '~') break;
        const size_t token_len = 4;
        if (cmd.compare(0, token_len, "GET ") == 0) {
            PXSGen *gen = ctrl.getitem((uint8_t *) (cmd.c_str() + token_len), (int) (cmd.size() - token_len));
            if (gen != NULL) {
                char *res = gen->consume_repr();
                std::cout << res << std::endl;
                free(res);
            }
        } else if (cmd.compare(0, token_len, "SET ") == 0) {
            const size_t pos = cmd.find("::");
            if (pos != std::string::npos) {
                std::string k = cmd.substr(token_len, pos - token_len), v = cmd.substr(pos);
                ctrl.setitem((uint8_t *) k.c_str(), (int) k.size(), (uint8_t *) v.c_str(), (int) v.size());
                const int diff = (int) cmd.size() - ctrl.last_encoded_len();
                std::cout << "\xe6\x9c\xac\xe6\xac\xa1\xe8\x8a\x82\xe7\xba\xa6\xe5\x86\x85\xe5\xad\x98\xe6\x95\xb0 " << diff << std::endl;
                total_diff += diff;
                std::cout << "\xe6\x80\xbb\xe5\x85\xb1\xe8\x8a\x82\xe7\xba\xa6\xe5\x86\x85\xe5\xad\x98\xe6\x95\xb0 " << total_diff << std::endl;
            }
        }
    }
    ctrl.free_prop();
    return 0;
}

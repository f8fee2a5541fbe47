#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
extern "C" uint32_t hostemu_decode(const uint8_t *data, size_t size, uint8_t **out, int32_t *w, int32_t *h, int32_t *stride);
int main(int argc, char **argv) {
    for (int i = 1; i < argc; ++i) {
        FILE *f = fopen(argv[i], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<uint8_t> buf((size_t) n); fread(buf.data(), 1, (size_t) n, f); fclose(f);
        uint8_t *out = nullptr; int32_t w, h, s;
        uint32_t e = hostemu_decode(buf.data(), buf.size(), &out, &w, &h, &s);
        printf("%s: err %08x %dx%d\n", argv[i], e, w, h);
        free(out);
    }
}

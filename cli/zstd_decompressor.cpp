// zstd-decompressor -- command line front end of the B200 decode path, flag compatible with the reference's
// src/main.rs:7-60:
//
//   zstd-decompressor <FILENAME> [-i|--info] [-o|--output <filename>] [-p|--print-skippable]
//
//   default   decode every Zstandard frame, concatenate, print to stdout (main.rs:42-58); skippable payloads are part
//             of the output only with -p; the whole output is buffered and nothing is written if any frame fails
//             (main.rs:51); like the reference the output must be valid UTF-8 (main.rs:55,57: from_utf8().unwrap()),
//             unless --binary is given (an extension: raw bytes out)
//   -i        dump the parsed frames in the layout of Rust's `{:#x?}` (main.rs:35-40) and exit.  Frame, Skippable, Header,
//             RawBlock and RLEBlock print exactly as the derived Debug impls do; a CompressedBlock prints its size only
//             (the reference dumps its parsed Huffman tree and FSE tables, which live on the GPU here).  Needs no GPU.
//   -o FILE   write to FILE (truncating) instead of stdout
//
// Decoding goes through the C ABI (include/zsb.h): zsb_scan on the host, zsb_decompress on the GPU; there is no CPU path.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../include/zsb.h"

static void usage(const char *argv0) {
    fprintf(stderr, "Usage: %s [OPTIONS] <FILENAME>\n\nArguments:\n  <FILENAME>  ZStandard file input, decompress it and output to stdout\n\n"
                    "Options:\n  -i, --info               Dump information about frames instead of outputing the result\n"
                    "  -o, --output <filename>  Output to given file (overwritting) instead of writing to stdout\n"
                    "  -p, --print-skippable    Output Skippable frames as well\n      --binary             Do not require the output to be valid UTF-8\n"
                    "  -h, --help               Print help\n", argv0);
}

static bool valid_utf8(const uint8_t *p, size_t n) {
    size_t i = 0;
    while (i < n) {
        const uint8_t c = p[i];
        size_t k; uint32_t cp, lo;
        if (c < 0x80) { i++; continue; }
        else if ((c & 0xE0) == 0xC0) { k = 1; cp = c & 0x1F; lo = 0x80; }
        else if ((c & 0xF0) == 0xE0) { k = 2; cp = c & 0x0F; lo = 0x800; }
        else if ((c & 0xF8) == 0xF0) { k = 3; cp = c & 0x07; lo = 0x10000; }
        else return false;
        if (i + k >= n) return false;
        for (size_t j = 1; j <= k; j++) { if ((p[i + j] & 0xC0) != 0x80) return false; cp = (cp << 6) | (p[i + j] & 0x3F); }
        if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
        i += k + 1;
    }
    return true;
}

// ---- `{:#x?}` of the frames (frame.rs:46-56,102-108,190-195; block.rs:28-40)
static void ind(std::string &o, int d) { o.append((size_t)d * 4, ' '); }
static void hex(std::string &o, uint64_t v) { char b[32]; snprintf(b, sizeof b, "0x%llx", (unsigned long long)v); o += b; }
static void opt(std::string &o, int d, const char *name, bool some, uint64_t v) {
    ind(o, d); o += name; o += ": ";
    if (!some) { o += "None,\n"; return; }
    o += "Some(\n"; ind(o, d + 1); hex(o, v); o += ",\n"; ind(o, d); o += "),\n";
}
static void bytes(std::string &o, int d, const uint8_t *p, size_t n) {
    if (n == 0) { o += "[]"; return; }
    o += "[\n";
    for (size_t i = 0; i < n; i++) { ind(o, d + 1); hex(o, p[i]); o += ",\n"; }
    ind(o, d); o += "]";
}
static std::string info(const uint8_t *src, const zsb_frame *fr, size_t nf, const zsb_block *bl) {
    std::string o;
    for (size_t f = 0; f < nf; f++) {
        const zsb_frame &F = fr[f];
        if (F.status) break;
        if (F.kind == 1) {
            const zsb_block &b = bl[F.first_block];
            o += "SkippableFrame(\n"; ind(o, 1); o += "Skippable {\n";
            ind(o, 2); o += "magic: "; hex(o, F.magic); o += ",\n";
            ind(o, 2); o += "data: "; bytes(o, 2, src + b.src_off, b.size); o += ",\n";
            ind(o, 1); o += "},\n)\n";
            continue;
        }
        o += "ZStandardFrame(\n"; ind(o, 1); o += "ZStandard {\n";
        ind(o, 2); o += "header: Header {\n";
        ind(o, 3); o += "content_checksum_flag: "; o += F.has_checksum ? "true" : "false"; o += ",\n";
        ind(o, 3); o += "window_size: "; hex(o, F.window_size); o += ",\n";
        opt(o, 3, "dictionnary_id", F.has_dict_id, F.dict_id);
        opt(o, 3, "content_size", F.has_content_size, F.content_size);
        ind(o, 2); o += "},\n";
        ind(o, 2); o += "blocks: [\n";
        for (uint32_t k = 0; k < F.n_blocks; k++) {
            const zsb_block &b = bl[F.first_block + k];
            if (b.type == 0) { ind(o, 3); o += "RawBlock(\n"; ind(o, 4); bytes(o, 4, src + b.src_off, b.size); o += ",\n"; ind(o, 3); o += "),\n"; }
            else if (b.type == 1) {
                ind(o, 3); o += "RLEBlock {\n"; ind(o, 4); o += "byte: "; hex(o, src[b.src_off]); o += ",\n";
                ind(o, 4); o += "repeat: "; hex(o, b.size); o += ",\n"; ind(o, 3); o += "},\n";
            } else {
                ind(o, 3); o += "CompressedBlock {\n"; ind(o, 4); o += "size: "; hex(o, b.size); o += ",\n"; ind(o, 3); o += "},\n";
            }
        }
        ind(o, 2); o += "],\n";
        opt(o, 2, "checksum", F.has_checksum, F.stored_checksum);
        ind(o, 1); o += "},\n)\n";
    }
    return o;
}

int main(int argc, char **argv) {
    const char *filename = nullptr, *output = nullptr;
    bool want_info = false, skippable = false, binary = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "-i" || a == "--info") want_info = true;
        else if (a == "-p" || a == "--print-skippable") skippable = true;
        else if (a == "--binary") binary = true;
        else if (a == "-h" || a == "--help") { usage(argv[0]); return 0; }
        else if (a == "-o" || a == "--output") { if (++i >= argc) { usage(argv[0]); return 2; } output = argv[i]; }
        else if (a.rfind("--output=", 0) == 0) output = argv[i] + 9;
        else if (a.size() > 1 && a[0] == '-') { fprintf(stderr, "error: unexpected argument '%s' found\n", argv[i]); usage(argv[0]); return 2; }
        else if (!filename) filename = argv[i];
        else { fprintf(stderr, "error: unexpected argument '%s' found\n", argv[i]); return 2; }
    }
    if (!filename) { fprintf(stderr, "error: the following required arguments were not provided:\n  <FILENAME>\n"); usage(argv[0]); return 2; }
    FILE *fp = fopen(filename, "rb");
    if (!fp) { fprintf(stderr, "Error: %s: %s\n", filename, strerror(errno)); return 1; }
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16]; size_t got;
    while ((got = fread(buf, 1, sizeof buf, fp)) > 0) data.insert(data.end(), buf, buf + got);
    fclose(fp);
    const uint32_t flags = ZSB_REFERENCE_QUIRKS | ZSB_VERIFY_CHECKSUM | (skippable ? ZSB_PRINT_SKIPPABLE : 0u);

    if (want_info) {            // main.rs:34-41: print every frame until the first error
        zsb_frame *fr = nullptr; zsb_block *bl = nullptr; size_t nf = 0, nb = 0; uint64_t ea = 0, eb = 0;
        const int rc = zsb_scan(data.data(), data.size(), flags, 0, &fr, &nf, &bl, &nb, &ea, &eb);
        const std::string o = info(data.data(), fr, nf, bl);
        fwrite(o.data(), 1, o.size(), stdout);
        zsb_free(fr); zsb_free(bl);
        if (rc) { fprintf(stderr, "Error: %s\n", zsb_strerror(rc)); return 1; }
        return 0;
    }

    zsb_ctx *ctx = nullptr;
    if (zsb_ctx_create(&ctx, 0) != ZSB_OK) { fprintf(stderr, "Error: no usable CUDA device (this decoder has no CPU path)\n"); return 1; }
    uint8_t *out = nullptr; size_t n = 0; uint64_t ea = 0, eb = 0;
    const int rc = zsb_decompress(ctx, data.data(), data.size(), flags, &out, &n, &ea, &eb);
    if (rc) { fprintf(stderr, "Error: %s\n", zsb_strerror(rc)); zsb_ctx_destroy(ctx); return 1; }       // no partial output (main.rs:51)
    if (!binary && !valid_utf8(out, n)) {
        fprintf(stderr, "thread 'main' panicked: called `Result::unwrap()` on an `Err` value: FromUtf8Error (the output is not UTF-8; use --binary)\n");
        free(out); zsb_ctx_destroy(ctx); return 101;
    }
    FILE *of = output ? fopen(output, "wb") : stdout;
    if (!of) { fprintf(stderr, "Error: %s: %s\n", output, strerror(errno)); free(out); zsb_ctx_destroy(ctx); return 1; }
    fwrite(out, 1, n, of);
    if (output) fclose(of); else fflush(stdout);
    zsb_free(out);
    zsb_ctx_destroy(ctx);
    return 0;
}

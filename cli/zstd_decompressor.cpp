// zstd-decompressor -- command line front end of the B200 decode path, flag compatible with the reference's
// src/main.rs:7-60:
//
//   zstd-decompressor <FILENAME> [-i|--info] [-o|--output <filename>] [-p|--print-skippable]
//
//   default   decode every Zstandard frame, concatenate, print to stdout (main.rs:42-58); skippable payloads are part
//             of the output only with -p; the whole output is buffered and nothing is written if any frame fails
//             (main.rs:51); like the reference the output must be valid UTF-8 (main.rs:55,57: from_utf8().unwrap()),
//             unless --binary is given (an extension: raw bytes out)
//   -i        dump the parsed frames in the layout of Rust's `{:#x?}` (main.rs:35-40) and exit: Frame, Skippable, Header, the blocks and,
//             for a CompressedBlock, its LiteralsSection (the Huffman tree through the reference's own Debug impl, huffman.rs:60-77:
//             one field per leaf, named " <code bits>", left first) and its Sequences (FSE tables state by state, fse.rs:72-89), as
//             the derived Debug impls print them.  Parsed on the host (zsb_scan, zsb_block_sections): needs no GPU, decodes nothing.
//   -o FILE   write to FILE (truncating) instead of stdout
//
// Decoding goes through the C ABI (include/zsb.h): zsb_scan on the host, zsb_decompress on the GPU; there is no CPU path.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <utility>
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
static void field_hex(std::string &o, int d, const char *name, uint64_t v) { ind(o, d); o += name; o += ": "; hex(o, v); o += ",\n"; }
static void mode(std::string &o, int d, const char *name, const zsb_sections &S, int t) {
    ind(o, d); o += name; o += ": ";
    if (S.mode[t] == 0) o += "PredefinedMode,\n";
    else if (S.mode[t] == 3) o += "RepeatMode,\n";
    else if (S.mode[t] == 1) { o += "RLEMode(\n"; ind(o, d + 1); hex(o, S.rle_symbol[t]); o += ",\n"; ind(o, d); o += "),\n"; }
    else {
        o += "FseCompressedMode(\n"; ind(o, d + 1); o += "FseTable {\n"; ind(o, d + 2); o += "table: [\n";
        for (uint32_t i = 0; i < (1u << S.al[t]); i++) {
            ind(o, d + 3); o += "State {\n";
            field_hex(o, d + 4, "output", S.table[t][i].output); field_hex(o, d + 4, "baseline", S.table[t][i].baseline); field_hex(o, d + 4, "bits_to_read", S.table[t][i].bits_to_read);
            ind(o, d + 3); o += "},\n";
        }
        ind(o, d + 2); o += "],\n"; field_hex(o, d + 2, "al", S.al[t]); ind(o, d + 1); o += "},\n"; ind(o, d); o += "),\n";
    }
}
// CompressedBlock { literals_section, sequences_section } (block.rs:36-39) at depth d
static void compressed_block(std::string &o, int d, const uint8_t *src, const zsb_sections &S) {
    ind(o, d); o += "CompressedBlock {\n";
    ind(o, d + 1); o += "literals_section: ";
    if (S.lit_type == 0) { o += "RawLiteralsBlock {\n"; ind(o, d + 2); o += "data: "; bytes(o, d + 2, src + S.lit_data_off, S.lit_data_len); o += ",\n"; ind(o, d + 1); o += "},\n"; }
    else if (S.lit_type == 1) { o += "RLELiteralsBlock {\n"; field_hex(o, d + 2, "byte", S.rle_byte); field_hex(o, d + 2, "repeat", S.regenerated_size); ind(o, d + 1); o += "},\n"; }
    else {
        o += "CompressedLiteralsBlock {\n";
        ind(o, d + 2); o += "huffman_decoder: ";
        if (S.lit_type == 3) o += "None,\n";
        else {
            o += "Some(\n"; ind(o, d + 3); o += "HuffmanDecoder {\n";
            // huffman.rs:60-77: a depth-first walk, left (0) before right (1): the leaves in the order of their codes read as bit strings
            std::vector<std::pair<std::string, int>> leaves;
            for (int sy = 0; sy < 256; sy++) if (S.code_len[sy]) {
                std::string c; for (int b = S.code_len[sy] - 1; b >= 0; b--) c += ((S.code[sy] >> b) & 1) ? '1' : '0';
                leaves.push_back({c, sy});
            }
            std::sort(leaves.begin(), leaves.end());
            for (auto &l : leaves) { ind(o, d + 4); o += " "; o += l.first; o += ": "; hex(o, (uint64_t)l.second); o += ",\n"; }
            ind(o, d + 3); o += "},\n"; ind(o, d + 2); o += "),\n";
        }
        field_hex(o, d + 2, "regenerated_size", S.regenerated_size);
        ind(o, d + 2); o += "jump_table: [\n"; for (int k = 0; k < 4; k++) { ind(o, d + 3); hex(o, S.jump_table[k]); o += ",\n"; } ind(o, d + 2); o += "],\n";
        ind(o, d + 2); o += "data: "; bytes(o, d + 2, src + S.lit_data_off, S.lit_data_len); o += ",\n";
        ind(o, d + 1); o += "},\n";
    }
    ind(o, d + 1); o += "sequences_section: Sequences {\n";
    field_hex(o, d + 2, "number_of_sequences", S.number_of_sequences);
    mode(o, d + 2, "literal_lengths_mode", S, 0); mode(o, d + 2, "offsets_mode", S, 1); mode(o, d + 2, "match_lengths_mode", S, 2);
    ind(o, d + 2); o += "bitstream: "; bytes(o, d + 2, src + S.bitstream_off, S.bitstream_len); o += ",\n";
    ind(o, d + 1); o += "},\n";
    ind(o, d); o += "},\n";
}
// -> the dump of every frame in front of the first error; rc = that error (a frame's section errors are raised while it is parsed,
// frame.rs:210-223: such a frame prints nothing)
static std::string info(const uint8_t *src, size_t n, const zsb_frame *fr, size_t nf, const zsb_block *bl, uint32_t flags, int &rc) {
    std::string o;
    static zsb_sections S;
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
        std::string z;
        z += "ZStandardFrame(\n"; ind(z, 1); z += "ZStandard {\n";
        ind(z, 2); z += "header: Header {\n";
        ind(z, 3); z += "content_checksum_flag: "; z += F.has_checksum ? "true" : "false"; z += ",\n";
        ind(z, 3); z += "window_size: "; hex(z, F.window_size); z += ",\n";
        opt(z, 3, "dictionnary_id", F.has_dict_id, F.dict_id);
        opt(z, 3, "content_size", F.has_content_size, F.content_size);
        ind(z, 2); z += "},\n";
        ind(z, 2); z += "blocks: [\n";
        for (uint32_t k = 0; k < F.n_blocks; k++) {
            const zsb_block &b = bl[F.first_block + k];
            if (b.type == 0) { ind(z, 3); z += "RawBlock(\n"; ind(z, 4); bytes(z, 4, src + b.src_off, b.size); z += ",\n"; ind(z, 3); z += "),\n"; }
            else if (b.type == 1) {
                ind(z, 3); z += "RLEBlock {\n"; ind(z, 4); z += "byte: "; hex(z, src[b.src_off]); z += ",\n";
                ind(z, 4); z += "repeat: "; hex(z, b.size); z += ",\n"; ind(z, 3); z += "},\n";
            } else {
                const int brc = zsb_block_sections(src, n, &b, flags, &S);
                if (brc) { rc = brc; return o; }
                compressed_block(z, 3, src, S);
            }
        }
        ind(z, 2); z += "],\n";
        opt(z, 2, "checksum", F.has_checksum, F.stored_checksum);
        ind(z, 1); z += "},\n)\n";
        o += z;
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
        int rc = zsb_scan(data.data(), data.size(), flags, 0, &fr, &nf, &bl, &nb, &ea, &eb);
        int src_rc = 0;
        const std::string o = info(data.data(), data.size(), fr, nf, bl, flags, src_rc);
        if (src_rc) rc = src_rc;
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

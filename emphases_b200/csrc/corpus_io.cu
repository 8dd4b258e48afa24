// Native corpus ingest / egress for emphases.from_files_to_files (host code).
//
// Replaces, for a whole file list at once and on a thread pool, the per-file
// Python work of the reference loop (emphases/core.py:169-179):
//   emphases.load.audio      torchaudio.load of a wav        emphases/load.py:11-17
//   pypar.Alignment(file)    Praat TextGrid parse             emphases/core.py:49
//   alignment.save(...)      TextGrid write                   emphases/core.py:111
// 16-bit PCM samples go straight into a caller-provided (pinned) int16 buffer
// -- converted to float on the GPU, exactly x / 32768 like torchaudio -- and
// word times into a float64 array parsed with strtod (= Python float()).
// Files this reader does not understand (other encodings, UTF-16 TextGrids)
// get a non-zero status and are left to the Python path.
#include <atomic>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/emphases_b200.h"

namespace {

constexpr const char* kSilence = "<silent>";

struct Word {
    std::string label;
    double start, end;
};

struct FileEntry {
    int status = 0;                 // 0 ok
    std::string error;
    std::string audio_path;
    int sample_rate = 0, channels = 0, bits = 0, format = 0;
    long long n_samples = 0;        // per channel
    long long data_offset = 0;      // byte offset of the PCM data in the file
    std::vector<Word> words;
};

bool read_file(const std::string& path, std::string& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long size = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(size > 0 ? size : 0);
    size_t got = size > 0 ? std::fread(&out[0], 1, size, f) : 0;
    std::fclose(f);
    return got == (size_t)(size > 0 ? size : 0);
}

void fail(FileEntry& e, int status, const std::string& message) {
    if (e.status == 0) {
        e.status = status;
        e.error = message;
    }
}

// ---- wav header ----
void scan_wav(FileEntry& e) {
    FILE* f = std::fopen(e.audio_path.c_str(), "rb");
    if (!f) return fail(e, 1, "cannot open " + e.audio_path);
    unsigned char head[12];
    if (std::fread(head, 1, 12, f) != 12 || std::memcmp(head, "RIFF", 4) || std::memcmp(head + 8, "WAVE", 4)) {
        std::fclose(f);
        return fail(e, 2, "not a RIFF/WAVE file: " + e.audio_path);
    }
    bool have_fmt = false, have_data = false;
    long long cursor = 12;
    unsigned char chunk[8];
    while (std::fread(chunk, 1, 8, f) == 8) {
        unsigned size = chunk[4] | (chunk[5] << 8) | (chunk[6] << 16) | ((unsigned)chunk[7] << 24);
        cursor += 8;
        if (!std::memcmp(chunk, "fmt ", 4)) {
            unsigned char fmt[40] = {0};
            unsigned want = size < 40 ? size : 40;
            if (std::fread(fmt, 1, want, f) != want) break;
            e.format = fmt[0] | (fmt[1] << 8);
            e.channels = fmt[2] | (fmt[3] << 8);
            e.sample_rate = fmt[4] | (fmt[5] << 8) | (fmt[6] << 16) | ((unsigned)fmt[7] << 24);
            e.bits = fmt[14] | (fmt[15] << 8);
            if (e.format == 0xFFFE && size >= 26) e.format = fmt[24] | (fmt[25] << 8);
            have_fmt = true;
            std::fseek(f, cursor + size + (size & 1), SEEK_SET);
        } else if (!std::memcmp(chunk, "data", 4)) {
            e.data_offset = cursor;
            if (have_fmt && e.channels > 0 && e.bits > 0)
                e.n_samples = (long long)size / (e.channels * (e.bits / 8));
            have_data = true;
            break;
        } else {
            std::fseek(f, cursor + size + (size & 1), SEEK_SET);
        }
        cursor += size + (size & 1);
    }
    std::fclose(f);
    if (!have_fmt || !have_data) return fail(e, 2, "missing fmt/data chunk: " + e.audio_path);
    if (e.format != 1 || e.bits != 16) return fail(e, 3, "not 16-bit PCM: " + e.audio_path);
}

// ---- TextGrid ----
struct Token {
    bool is_string;
    std::string text;
    double number;
};

void tokenize(const std::string& s, std::vector<Token>& tokens) {
    size_t i = 0, n = s.size();
    while (i < n) {
        char c = s[i];
        if (c == '"') {
            std::string text;
            ++i;
            while (i < n) {
                if (s[i] == '"') {
                    if (i + 1 < n && s[i + 1] == '"') { text.push_back('"'); i += 2; continue; }
                    break;
                }
                text.push_back(s[i++]);
            }
            ++i;
            tokens.push_back({true, text, 0.0});
        } else if (c == '[') {                       // "intervals [3]:" index
            while (i < n && s[i] != ']') ++i;
            ++i;
        } else if ((c >= '0' && c <= '9') || ((c == '-' || c == '+') && i + 1 < n && s[i + 1] >= '0' && s[i + 1] <= '9')) {
            char* end = nullptr;
            double v = std::strtod(s.c_str() + i, &end);
            size_t used = end - (s.c_str() + i);
            if (used == 0) { ++i; continue; }
            tokens.push_back({false, std::string(), v});
            i += used;
        } else {
            ++i;
        }
    }
}

void parse_textgrid(const std::string& path, FileEntry& e) {
    std::string raw;
    if (!read_file(path, raw)) return fail(e, 4, "cannot open " + path);
    if (raw.size() >= 2 && ((unsigned char)raw[0] == 0xFF || (unsigned char)raw[0] == 0xFE) &&
        ((unsigned char)raw[1] == 0xFE || (unsigned char)raw[1] == 0xFF))
        return fail(e, 5, "UTF-16 TextGrid: " + path);
    if (raw.size() >= 3 && (unsigned char)raw[0] == 0xEF) raw.erase(0, 3);   // UTF-8 BOM
    std::vector<Token> tokens;
    tokenize(raw, tokens);
    std::vector<Word> first, chosen;
    bool have_first = false, have_chosen = false;
    for (size_t i = 0; i + 4 < tokens.size(); ++i) {
        if (!tokens[i].is_string) continue;
        const bool interval = tokens[i].text == "IntervalTier";
        if (!interval && tokens[i].text != "TextTier") continue;
        if (!tokens[i + 1].is_string || tokens[i + 4].is_string) continue;
        const std::string name = tokens[i + 1].text;
        const long count = (long)tokens[i + 4].number;
        size_t cursor = i + 5;
        if (!interval) { i = cursor + 2 * count - 1; continue; }
        std::vector<Word> words;
        bool ok = true;
        for (long j = 0; j < count; ++j, cursor += 3) {
            if (cursor + 2 >= tokens.size()) { ok = false; break; }
            if (tokens[cursor].is_string || tokens[cursor + 1].is_string || !tokens[cursor + 2].is_string) { ok = false; break; }
            words.push_back({tokens[cursor + 2].text, tokens[cursor].number, tokens[cursor + 1].number});
        }
        if (!ok) return fail(e, 6, "malformed TextGrid: " + path);
        std::string lower = name;
        for (auto& ch : lower) ch = (char)std::tolower((unsigned char)ch);
        if (!have_first) { first = words; have_first = true; }
        if (!have_chosen && (lower == "words" || lower == "word")) { chosen = words; have_chosen = true; }
        i = cursor - 1;
    }
    if (!have_first) return fail(e, 6, "no interval tier: " + path);
    const std::vector<Word>& source = have_chosen ? chosen : first;
    // fill gaps with silences so the words tile [0, end] (alignment.py:_fill_gaps)
    double at = 0.0;
    for (const Word& w : source) {
        std::string label = w.label;
        size_t a = label.find_first_not_of(" \t\r\n"), b = label.find_last_not_of(" \t\r\n");
        std::string trimmed = a == std::string::npos ? "" : label.substr(a, b - a + 1);
        if (trimmed.empty() || trimmed == "sp" || trimmed == "sil") label = kSilence;
        if (w.start - at > 1e-9) e.words.push_back({kSilence, at, w.start});
        e.words.push_back({label, w.start, w.end});
        at = w.end;
    }
}

std::string format_double(double v) {
    char buffer[64];
    auto result = std::to_chars(buffer, buffer + sizeof(buffer), v);
    std::string s(buffer, result.ptr);
    if (s.find_first_of(".en") == std::string::npos) s += ".0";
    return s;
}

bool write_textgrid(const std::string& path, const std::vector<Word>& words) {
    std::string out;
    const std::string xmax = format_double(words.empty() ? 0.0 : words.back().end);
    out += "File type = \"ooTextFile\"\nObject class = \"TextGrid\"\n\nxmin = 0\nxmax = " + xmax +
           "\ntiers? <exists>\nsize = 2\nitem []:\n";
    const char* names[2] = {"words", "phones"};
    for (int tier = 0; tier < 2; ++tier) {
        out += "    item [" + std::to_string(tier + 1) + "]:\n        class = \"IntervalTier\"\n        name = \"" +
               names[tier] + "\"\n        xmin = 0\n        xmax = " + xmax +
               "\n        intervals: size = " + std::to_string(words.size()) + "\n";
        for (size_t j = 0; j < words.size(); ++j) {
            std::string label;
            if (words[j].label == kSilence) {
                label = tier == 0 ? "sp" : "sil";
            } else {
                for (char c : words[j].label) { if (c == '"') label += "\"\""; else label.push_back(c); }
            }
            out += "        intervals [" + std::to_string(j + 1) + "]:\n            xmin = " +
                   format_double(words[j].start) + "\n            xmax = " + format_double(words[j].end) +
                   "\n            text = \"" + label + "\"\n";
        }
    }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    size_t wrote = std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    return wrote == out.size();
}

template <typename Fn>
void parallel_for(int n, int n_threads, Fn fn) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n) n_threads = n > 0 ? n : 1;
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&]() {
            for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
        });
    for (auto& t : pool) t.join();
}

}  // namespace

struct emph_corpus {
    std::vector<FileEntry> files;
};

extern "C" {

emph_corpus* emph_corpus_open(
    const char* const* text_paths, const char* const* audio_paths, int32_t n_files, int32_t n_threads) {
    emph_corpus* corpus = new emph_corpus();
    corpus->files.resize(n_files > 0 ? n_files : 0);
    parallel_for(n_files, n_threads, [&](int i) {
        FileEntry& e = corpus->files[i];
        e.audio_path = audio_paths[i];
        scan_wav(e);
        parse_textgrid(text_paths[i], e);
    });
    return corpus;
}

int emph_corpus_info(
    const emph_corpus* corpus, int32_t* status, int32_t* sample_rate, int32_t* channels,
    int64_t* n_samples, int32_t* n_words) {
    if (!corpus) return EMPH_EINVAL;
    for (size_t i = 0; i < corpus->files.size(); ++i) {
        const FileEntry& e = corpus->files[i];
        status[i] = e.status;
        sample_rate[i] = e.sample_rate;
        channels[i] = e.channels;
        n_samples[i] = e.n_samples;
        n_words[i] = (int32_t)e.words.size();
    }
    return EMPH_OK;
}

const char* emph_corpus_error(const emph_corpus* corpus, int32_t index) {
    if (!corpus || index < 0 || (size_t)index >= corpus->files.size()) return "";
    return corpus->files[index].error.c_str();
}

int emph_corpus_fill(
    emph_corpus* corpus, int16_t* audio_dst, const int64_t* sample_offsets,
    double* times_dst, const int64_t* word_offsets, int32_t n_threads) {
    if (!corpus) return EMPH_EINVAL;
    std::atomic<int> failures(0);
    parallel_for((int)corpus->files.size(), n_threads, [&](int i) {
        FileEntry& e = corpus->files[i];
        if (e.status != 0) return;
        double* times = times_dst + 2 * word_offsets[i];
        for (size_t j = 0; j < e.words.size(); ++j) {
            times[2 * j] = e.words[j].start;
            times[2 * j + 1] = e.words[j].end;
        }
        FILE* f = std::fopen(e.audio_path.c_str(), "rb");
        if (!f) { fail(e, 1, "cannot reopen " + e.audio_path); ++failures; return; }
        std::fseek(f, (long)e.data_offset, SEEK_SET);
        int16_t* dst = audio_dst + sample_offsets[i];
        bool ok = true;
        if (e.channels == 1) {
            ok = std::fread(dst, 2, (size_t)e.n_samples, f) == (size_t)e.n_samples;
        } else {                                     // keep channel 0 (mels.py:48)
            std::vector<int16_t> frame_buffer((size_t)e.channels * 4096);
            long long done = 0;
            while (done < e.n_samples && ok) {
                long long want = e.n_samples - done < 4096 ? e.n_samples - done : 4096;
                ok = std::fread(frame_buffer.data(), 2 * e.channels, (size_t)want, f) == (size_t)want;
                for (long long k = 0; k < want; ++k) dst[done + k] = frame_buffer[(size_t)k * e.channels];
                done += want;
            }
        }
        std::fclose(f);
        if (!ok) { fail(e, 1, "short read: " + e.audio_path); ++failures; }
    });
    return failures.load() == 0 ? EMPH_OK : EMPH_EINVAL;
}

int emph_corpus_write_textgrids(
    const emph_corpus* corpus, const char* const* output_paths, int32_t n_threads) {
    if (!corpus) return EMPH_EINVAL;
    std::atomic<int> failures(0);
    parallel_for((int)corpus->files.size(), n_threads, [&](int i) {
        const FileEntry& e = corpus->files[i];
        if (e.status != 0 || output_paths[i] == nullptr || output_paths[i][0] == 0) return;
        if (!write_textgrid(output_paths[i], e.words)) ++failures;
    });
    return failures.load() == 0 ? EMPH_OK : EMPH_EINVAL;
}

void emph_corpus_close(emph_corpus* corpus) { delete corpus; }

}  // extern "C"

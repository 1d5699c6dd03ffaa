/*
 * pthost_formats.c — data formats on either side of the hot path:
 *   pth_load_pam     the reader matching the reference's writer: pamalign.h:166-210 (+ header parser :52-131)
 *   pth_save_ppm     a real binary P6 PPM (the reference's "result.ppm" is a PAM file with a .ppm name)
 *   pth_save_png     8-bit RGBA PNG, self-contained (zlib "stored" blocks, CRC-32 and Adler-32 computed here)
 *   pth_import_obj   Wavefront OBJ -> triangles.txt in the layout parseTrianglesFromFile expects
 *                    (CLSuperPathTracer.c:77-118): x\ny\nz\n\n per vertex, one more \n between triangles, NO
 *                    trailing newline (a complete 13-line tail would make the feof() loop read a spurious triangle)
 * Plain C, host only.
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pthost.h"

/* ------------------------------------------------------------------------------------------ PAM reader */
#define PAM_IS_NL(x) ((x) == 0x0A || (x) == 0x0D)
#define PAM_IS_SPACE(x) ((x) == 0x20 || (x) == 0x09 || PAM_IS_NL(x))

/* pamalign.h:52-131.  One header line = "<label><space><value>\n"; unknown labels are skipped; TUPLTYPE is ignored. */
static int pam_read_header(FILE *fp, pth_image *img) {
    int seen_w = 0, seen_h = 0, seen_d = 0, seen_m = 0, seen_end = 0;
    enum { BUFLEN = 256 };
    while (!seen_end) {
        char buffer[BUFLEN], label[BUFLEN + 1], value[BUFLEN + 1];
        if (!fgets(buffer, sizeof buffer, fp)) {
            fprintf(stderr, "EOF or error reading file while trying to read the PAM header\n");
            return 1;
        }
        int invalue = 0, cursor = 0;
        value[0] = '\0';
        for (char *rc = buffer; rc < buffer + BUFLEN; ++rc) {
            if (invalue) {
                value[cursor++] = *rc;
                if (PAM_IS_NL(*rc) || *rc == '\0') { value[--cursor] = '\0'; break; }
            } else {
                label[cursor++] = *rc;
                if (PAM_IS_SPACE(*rc)) { invalue = 1; label[--cursor] = '\0'; cursor = 0; }
                if (PAM_IS_NL(*rc) || *rc == '\0') break;
            }
        }
        if (!invalue) continue;
        if (!strcmp(label, "ENDHDR")) seen_end = 1;
        else if (!strcmp(label, "WIDTH")) { seen_w = 1; img->width = (uint32_t)atoi(value); }
        else if (!strcmp(label, "HEIGHT")) { seen_h = 1; img->height = (uint32_t)atoi(value); }
        else if (!strcmp(label, "DEPTH")) { seen_d = 1; img->channels = (uint32_t)atoi(value); }
        else if (!strcmp(label, "MAXVAL")) {
            seen_m = 1;
            img->maxval = (uint32_t)atoi(value);
            if (img->maxval <= 0xff) img->depth = 8;
            else if (img->maxval <= 0xffff) img->depth = 16;
            else { fprintf(stderr, "maxval too high\n"); return 1; }
        }
    }
    if (!seen_w || !seen_h || !seen_d || !seen_m) { fprintf(stderr, "incomplete header\n"); return 1; }
    return 0;
}

/* pamalign.h:166-210.  3-channel images are padded to 4 values per pixel (the pad value is left at 0 here; the
 * reference leaves it uninitialised); 16-bit samples are big-endian in the file, host-endian in memory. */
int pth_load_pam(const char *path, pth_image *img) {
    memset(img, 0, sizeof(*img));
    FILE *fp = fopen(path, "rb");
    if (!fp) { fprintf(stderr, "could not open %s\n", path); return 1; }
    char hdr[4] = {0, 0, 0, 0};
    if (fread(hdr, 3, 1, fp) != 1 || strcmp(hdr, "P7\n")) {
        fprintf(stderr, "not a PAM file: %s\n", path);
        fclose(fp);
        return 1;
    }
    if (pam_read_header(fp, img)) { fclose(fp); return 1; }      /* the reference ignores this status and carries on */
    if (img->channels < 1 || img->channels > 4) {
        fprintf(stderr, "can't process PAM file with %u channels\n", img->channels);
        fclose(fp);
        return 1;
    }
    const uint32_t stride = img->channels + (img->channels == 3);
    size_t bytes = (size_t)(img->depth / 8) * stride * img->width * img->height;
    img->data_size = bytes;
    img->data = calloc(bytes ? bytes : 1, 1);
    if (!img->data) { fprintf(stderr, "can't allocate memory for image data\n"); fclose(fp); return 1; }
    uint8_t *d8 = (uint8_t *)img->data;
    uint16_t *d16 = (uint16_t *)img->data;
    size_t cur = 0;
    const size_t npix = (size_t)img->width * img->height;
    int status = 0;
    for (size_t p = 0; p < npix && !status; ++p) {
        for (uint32_t ch = 0; ch < img->channels; ++ch, ++cur) {
            int a = fgetc(fp);
            if (a == EOF) { status = 1; break; }
            if (img->depth == 8) d8[cur] = (uint8_t)a;
            else {
                int b = fgetc(fp);
                if (b == EOF) { status = 1; break; }
                d16[cur] = (uint16_t)((a << 8) | b);
            }
        }
        if (img->channels == 3) ++cur;
    }
    fclose(fp);
    if (status) fprintf(stderr, "truncated PAM file: %s\n", path);
    return status;
}

void pth_free_image(pth_image *img) {
    if (img) { free(img->data); img->data = NULL; img->data_size = 0; }
}

/* ------------------------------------------------------------------------------------------ P6 PPM */
int pth_save_ppm(const char *path, int width, int height, const void *rgba8) {
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "could not open %s for writing\n", path); return 1; }
    fprintf(f, "P6\n%d %d\n255\n", width, height);
    const uint8_t *src = (const uint8_t *)rgba8;
    uint8_t *row = (uint8_t *)malloc((size_t)width * 3 + 1);
    int bad = row == NULL;
    for (int y = 0; y < height && !bad; ++y) {
        for (int x = 0; x < width; ++x) {
            const uint8_t *p = src + 4 * ((size_t)y * width + x);
            row[3 * x] = p[0]; row[3 * x + 1] = p[1]; row[3 * x + 2] = p[2];
        }
        bad = fwrite(row, 3, (size_t)width, f) != (size_t)width;
    }
    free(row);
    fclose(f);
    return bad;
}

/* ------------------------------------------------------------------------------------------ PNG */
static uint32_t crc_table[256];
static void crc_init(void) {
    if (crc_table[1]) return;
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        crc_table[n] = c;
    }
}
static uint32_t crc_update(uint32_t c, const uint8_t *p, size_t n) {
    for (size_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c;
}
static void put_be32(uint8_t *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

static int png_chunk(FILE *f, const char type[4], const uint8_t *data, size_t len) {
    uint8_t hdr[8];
    put_be32(hdr, (uint32_t)len);
    memcpy(hdr + 4, type, 4);
    uint32_t c = crc_update(0xFFFFFFFFu, hdr + 4, 4);
    if (len) c = crc_update(c, data, len);
    uint8_t tail[4];
    put_be32(tail, c ^ 0xFFFFFFFFu);
    return fwrite(hdr, 1, 8, f) != 8 || (len && fwrite(data, 1, len, f) != len) || fwrite(tail, 1, 4, f) != 4;
}

/* RGBA8, filter 0 on every scanline, deflate "stored" blocks (<= 65535 bytes each): bigger than a compressed
 * PNG but needs no zlib, and every decoder reads it. */
int pth_save_png(const char *path, int width, int height, const void *rgba8) {
    if (width <= 0 || height <= 0) return 1;
    crc_init();
    const size_t stride = (size_t)width * 4 + 1, raw = stride * (size_t)height;
    const size_t nblocks = (raw + 65534) / 65535;
    const size_t zlen = 2 + raw + 5 * nblocks + 4;
    uint8_t *z = (uint8_t *)malloc(zlen);
    if (!z) return 1;
    size_t o = 0;
    z[o++] = 0x78; z[o++] = 0x01;                                 /* zlib header: deflate, 32 K window, no preset */
    uint32_t a = 1, b = 0;                                        /* Adler-32 of the uncompressed stream */
    const uint8_t *src = (const uint8_t *)rgba8;
    size_t produced = 0, in_block = 0;
    for (int y = 0; y < height; ++y) {
        for (size_t k = 0; k < stride; ++k) {
            if (in_block == 0) {
                size_t n = raw - produced < 65535 ? raw - produced : 65535;
                z[o++] = (raw - produced <= 65535) ? 1 : 0;       /* BFINAL on the last block, BTYPE = 00 */
                z[o++] = n & 0xff; z[o++] = n >> 8; z[o++] = ~n & 0xff; z[o++] = (~n >> 8) & 0xff;
                in_block = n;
            }
            uint8_t v = k == 0 ? 0 : src[(size_t)y * width * 4 + (k - 1)];
            z[o++] = v;
            a += v; if (a >= 65521) a -= 65521;
            b += a; if (b >= 65521) b -= 65521;
            ++produced; --in_block;
        }
    }
    put_be32(z + o, (b << 16) | a);
    o += 4;
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "could not open %s for writing\n", path); free(z); return 1; }
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t ihdr[13];
    put_be32(ihdr, (uint32_t)width); put_be32(ihdr + 4, (uint32_t)height);
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;   /* 8 bit, colour type 6 (RGBA) */
    int bad = fwrite(sig, 1, 8, f) != 8 || png_chunk(f, "IHDR", ihdr, 13) || png_chunk(f, "IDAT", z, o) || png_chunk(f, "IEND", NULL, 0);
    fclose(f);
    free(z);
    return bad;
}

/* ------------------------------------------------------------------------------------------ OBJ importer */
/* v x y z [w] / f i j k ...  (i, i/t, i/t/n, i//n; negative = relative to the vertices read so far).  Polygons are
 * fan-triangulated.  Every coordinate becomes scale*c + translate[axis], printed with %f like the shipped meshes.
 * Returns the number of triangles written, -1 on I/O or syntax errors. */
long pth_import_obj(const char *obj_path, const char *triangles_txt_path, float scale, const float translate[3]) {
    FILE *in = fopen(obj_path, "r");
    if (!in) { fprintf(stderr, "could not open %s\n", obj_path); return -1; }
    FILE *out = fopen(triangles_txt_path, "w");
    if (!out) { fprintf(stderr, "could not open %s for writing\n", triangles_txt_path); fclose(in); return -1; }
    const float tr[3] = {translate ? translate[0] : 0.f, translate ? translate[1] : 0.f, translate ? translate[2] : 0.f};
    size_t cap = 1024, nv = 0;
    float *v = (float *)malloc(cap * 3 * sizeof(float));
    long ntri = 0;
    int err = v == NULL;
    char *line = NULL;
    size_t linecap = 0;
    while (!err && getline(&line, &linecap, in) >= 0) {
        char *p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            float c[3];
            if (sscanf(p + 1, "%f %f %f", &c[0], &c[1], &c[2]) != 3) { err = 1; break; }
            if (nv == cap) {
                cap *= 2;
                float *nvp = (float *)realloc(v, cap * 3 * sizeof(float));
                if (!nvp) { err = 1; break; }
                v = nvp;
            }
            for (int a = 0; a < 3; ++a) v[3 * nv + a] = scale * c[a] + tr[a];
            ++nv;
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            long idx[64];
            int n = 0;
            p += 1;
            while (n < 64) {
                while (*p == ' ' || *p == '\t') ++p;
                if (*p == '\0' || *p == '\n' || *p == '\r' || *p == '#') break;
                char *end;
                long i = strtol(p, &end, 10);
                if (end == p) { err = 1; break; }
                if (i < 0) i = (long)nv + i; else i -= 1;
                if (i < 0 || (size_t)i >= nv) { err = 1; break; }
                idx[n++] = i;
                p = end;
                while (*p && !isspace((unsigned char)*p)) ++p;         /* skip /t/n */
            }
            if (err) break;
            for (int k = 1; k + 1 < n; ++k) {
                const long tri[3] = {idx[0], idx[k], idx[k + 1]};
                if (ntri) fputs("\n\n\n", out);
                for (int c = 0; c < 3; ++c) {
                    const float *q = v + 3 * tri[c];
                    fprintf(out, "%s%f\n%f\n%f", c ? "\n\n" : "", q[0], q[1], q[2]);
                }
                ++ntri;
            }
        }
    }
    free(line);
    free(v);
    fclose(in);
    if (fclose(out)) err = 1;
    if (err) { fprintf(stderr, "error importing %s\n", obj_path); return -1; }
    return ntri;
}

/* CPU ORACLE (test infrastructure) -- range coder.
 * Restates /root/reference/src/range_coder/mod.rs:48-117, decoder.rs, encoder.rs. */
#include "oracle.h"

#include <string.h>

/* mod.rs:48-68 */
#define UINT_BITS 8u
#define BITRES 3u
#define WINDOW_SIZE 32u
#define SYM_BITS 8u
#define CODE_BITS 32u
#define SYM_MAX ((1u << SYM_BITS) - 1u)
#define CODE_SHIFT (CODE_BITS - SYM_BITS - 1u)
#define CODE_TOP (1u << (CODE_BITS - 1u))
#define CODE_BOT (CODE_TOP >> SYM_BITS)
#define CODE_EXTRA ((CODE_BITS - 2u) % SYM_BITS + 1u)

static inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

/* math.rs:5-7 */
uint32_t orc_ilog(uint32_t x) { return x ? 32u - (uint32_t)__builtin_clz(x) : 0u; }

/* mod.rs:84-86 */
uint32_t orc_tell(uint32_t bits_total, uint32_t rng) { return bits_total - orc_ilog(rng); }

/* mod.rs:96-111 */
uint32_t orc_tell_frac(uint32_t bits_total, uint32_t rng)
{
    static const uint32_t correction[8] = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535};
    uint32_t bits = bits_total << BITRES;
    uint32_t l = orc_ilog(rng);
    uint32_t r = rng >> (l - 16);
    uint32_t b = (r >> 12) - 8;
    if (r > correction[b]) b += 1;
    l = (l << 3) + b;
    return bits - l;
}

/* mod.rs:114-117 */
uint32_t orc_laplace_freq1(uint32_t fs0, uint32_t decay)
{
    uint32_t ft = 32768u - 32u - fs0;
    return (ft * (16384u - decay)) >> 15;
}

/* mod.rs:530-534 (test helper get_start_freq) */
uint32_t orc_laplace_start_freq(uint32_t decay)
{
    uint32_t ft = 32768u - 33u;
    uint32_t fs = (ft * (16384u - decay)) / (16384u + decay);
    return fs + 1u;
}

/* ------------------------------------------------------------------ decoder.rs */

/* decoder.rs:86-94 */
static uint8_t dec_read_byte(orc_dec *d)
{
    if (d->offs < d->storage) return d->buffer[d->offs++];
    return 0;
}

/* decoder.rs:97-104 */
static uint8_t dec_read_byte_from_end(orc_dec *d)
{
    if (d->end_offs < d->storage) {
        d->end_offs += 1;
        return d->buffer[d->storage - d->end_offs];
    }
    return 0;
}

/* decoder.rs:108-122 */
static void dec_normalize(orc_dec *d)
{
    while (d->rng <= CODE_BOT) {
        d->bits_total += SYM_BITS;
        d->rng <<= SYM_BITS;
        uint32_t symbol = d->rem;
        d->rem = dec_read_byte(d);
        symbol = ((symbol << SYM_BITS) | d->rem) >> (SYM_BITS - CODE_EXTRA);
        d->val = ((d->val << SYM_BITS) + (SYM_MAX & ~symbol)) & (CODE_TOP - 1u);
    }
}

/* decoder.rs:50-78 */
void orc_dec_init(orc_dec *d, const uint8_t *buf, uint32_t len)
{
    memset(d, 0, sizeof(*d));
    d->buffer = buf;
    d->storage = len;
    d->bits_total = CODE_BITS + 1u - ((CODE_BITS - CODE_EXTRA) / SYM_BITS) * SYM_BITS;
    d->rng = 1u << CODE_EXTRA;
    d->rem = dec_read_byte(d);
    d->val = d->rng - 1u - ((uint32_t)d->rem >> (SYM_BITS - CODE_EXTRA));
    dec_normalize(d);
}

/* decoder.rs:81-83 */
void orc_dec_shrink_storage(orc_dec *d, uint32_t by) { d->storage -= by; }

/* decoder.rs:143-147 */
uint32_t orc_dec_decode(orc_dec *d, uint32_t ft)
{
    d->ext = d->rng / ft;
    uint32_t s = d->val / d->ext;
    return ft - umin(s + 1u, ft);
}

/* decoder.rs:150-154 */
uint32_t orc_dec_decode_bin(orc_dec *d, uint32_t bits)
{
    d->ext = d->rng >> bits;
    uint32_t s = d->val / d->ext;
    return (1u << bits) - umin(s + 1u, 1u << bits);
}

/* decoder.rs:172-181 */
void orc_dec_update(orc_dec *d, uint32_t fl, uint32_t fh, uint32_t ft)
{
    uint32_t s = d->ext * (ft - fh);
    d->val -= s;
    d->rng = fl > 0 ? d->ext * (fh - fl) : d->rng - s;
    dec_normalize(d);
}

/* decoder.rs:184-195 */
int orc_dec_bit_logp(orc_dec *d, uint32_t logp)
{
    uint32_t r = d->rng, v = d->val, s = r >> logp;
    int ret = v < s;
    if (!ret) d->val = v - s;
    d->rng = ret ? s : r - s;
    dec_normalize(d);
    return ret;
}

/* decoder.rs:210-232 */
uint32_t orc_dec_icdf(orc_dec *d, const uint8_t *icdf, uint32_t ftb)
{
    uint32_t s = d->rng, v = d->val, r = s >> ftb, t, ret = 0;
    for (;;) {
        t = s;
        s = r * icdf[ret];
        if (v >= s) break;
        ret += 1;
    }
    d->val = v - s;
    d->rng = t - s;
    dec_normalize(d);
    return ret;
}

/* decoder.rs:279-303 */
uint32_t orc_dec_bits(orc_dec *d, uint32_t bits)
{
    uint32_t window = d->end_window, available = d->end_bits;
    if (available < bits) {
        do {
            window |= (uint32_t)dec_read_byte_from_end(d) << available;
            available += SYM_BITS;
        } while (available <= WINDOW_SIZE - SYM_BITS);
    }
    uint32_t ret = window & ((1u << bits) - 1u);
    window >>= bits;
    available -= bits;
    d->end_window = window;
    d->end_bits = available;
    d->bits_total += bits;
    return ret;
}

/* decoder.rs:245-266 */
uint32_t orc_dec_uint(orc_dec *d, uint32_t ft)
{
    ft -= 1u;
    uint32_t ftb = orc_ilog(ft);
    if (ftb > UINT_BITS) {
        ftb -= UINT_BITS;
        uint32_t ft1 = (ft >> ftb) + 1u;
        uint32_t s = orc_dec_decode(d, ft1);
        orc_dec_update(d, s, s + 1u, ft1);
        uint32_t t = (s << ftb) | orc_dec_bits(d, ftb);
        if (t <= ft) return t;
        return ft; /* corrupt frame: saturate (decoder.rs:258-259) */
    }
    ft += 1u;
    uint32_t s = orc_dec_decode(d, ft);
    orc_dec_update(d, s, s + 1u, ft);
    return s;
}

/* decoder.rs:314-355 */
int32_t orc_dec_laplace(orc_dec *d, uint32_t fs, uint32_t decay)
{
    int32_t val = 0;
    uint32_t fm = orc_dec_decode_bin(d, 15);
    uint32_t fl = 0;
    if (fm >= fs) {
        val += 1;
        fl = fs;
        fs = orc_laplace_freq1(fs, decay) + 1u;
        while (fs != 0 && fm >= fl + 2u * fs) {
            fs *= 2u;
            fl += fs;
            fs = ((fs - 2u) * decay) >> 15;
            fs += 1u;
            val += 1;
        }
        if (fs <= 1u) {
            uint32_t di = (fm - fl) >> 1;
            val += (int32_t)di;
            fl += 2u * di;
        }
        if (fm < fl + fs) val = -val;
        else fl += fs;
    }
    orc_dec_update(d, fl, umin(fl + fs, 32768u), 32768u);
    return val;
}

uint32_t orc_dec_tell(const orc_dec *d) { return orc_tell(d->bits_total, d->rng); }
uint32_t orc_dec_tell_frac(const orc_dec *d) { return orc_tell_frac(d->bits_total, d->rng); }

/* ------------------------------------------------------------------ encoder.rs */

/* encoder.rs:50-69 */
void orc_enc_init(orc_enc *e, uint8_t *buf, uint32_t len)
{
    memset(e, 0, sizeof(*e));
    e->buffer = buf;
    e->buffer_len = len;
    e->storage = len;
    e->bits_total = CODE_BITS + 1u;
    e->rng = CODE_TOP;
    e->rem = -1;
}

/* encoder.rs:91-99 */
static int enc_write_byte(orc_enc *e, uint8_t v)
{
    if (e->offs + e->end_offs >= e->storage) return e->error = ORC_ERR_BUFFER_TOO_SMALL;
    e->buffer[e->offs++] = v;
    return 0;
}

/* encoder.rs:102-109 */
static int enc_write_byte_at_end(orc_enc *e, uint8_t v)
{
    if (e->offs + e->end_offs >= e->storage) return e->error = ORC_ERR_BUFFER_TOO_SMALL;
    e->end_offs += 1;
    e->buffer[e->storage - e->end_offs] = v;
    return 0;
}

/* encoder.rs:124-153 */
static int enc_carry_out(orc_enc *e, uint32_t c)
{
    if (c != SYM_MAX) {
        uint32_t carry = c >> SYM_BITS;
        if (e->rem >= 0) {
            if (enc_write_byte(e, (uint8_t)((uint32_t)e->rem + carry))) return e->error;
        }
        if (e->ext > 0) {
            uint8_t sym = (uint8_t)((SYM_MAX + carry) & SYM_MAX);
            do {
                if (enc_write_byte(e, sym)) return e->error;
                e->ext -= 1;
            } while (e->ext != 0);
        }
        e->rem = (int32_t)(c & SYM_MAX);
    } else {
        e->ext += 1;
    }
    return 0;
}

/* encoder.rs:157-168 */
static int enc_normalize(orc_enc *e)
{
    while (e->rng <= CODE_BOT) {
        if (enc_carry_out(e, e->val >> CODE_SHIFT)) return e->error;
        e->val = (e->val << SYM_BITS) & (CODE_TOP - 1u);
        e->rng <<= SYM_BITS;
        e->bits_total += SYM_BITS;
    }
    return 0;
}

/* encoder.rs:187-198 */
int orc_enc_encode(orc_enc *e, uint32_t fl, uint32_t fh, uint32_t ft)
{
    uint32_t r = e->rng / ft;
    if (fl > 0) {
        e->val += e->rng - (r * (ft - fl));
        e->rng = r * (fh - fl);
    } else {
        e->rng -= r * (ft - fh);
    }
    return enc_normalize(e);
}

/* encoder.rs:201-212 */
int orc_enc_encode_bin(orc_enc *e, uint32_t fl, uint32_t fh, uint32_t bits)
{
    uint32_t r = e->rng >> bits;
    if (fl > 0) {
        e->val += e->rng - (r * ((1u << bits) - fl));
        e->rng = r * (fh - fl);
    } else {
        e->rng -= r * ((1u << bits) - fh);
    }
    return enc_normalize(e);
}

/* encoder.rs:215-227 */
int orc_enc_bit_logp(orc_enc *e, uint32_t val, uint32_t logp)
{
    uint32_t r = e->rng, l = e->val, s = r >> logp;
    r -= s;
    if (val != 0) e->val = l + r;
    e->rng = val != 0 ? s : r;
    return enc_normalize(e);
}

/* encoder.rs:239-250 */
int orc_enc_icdf(orc_enc *e, uint32_t s, const uint8_t *icdf, uint32_t ftb)
{
    uint32_t r = e->rng >> ftb;
    if (s > 0) {
        e->val += e->rng - (r * icdf[s - 1]);
        e->rng = r * (uint32_t)(uint8_t)(icdf[s - 1] - icdf[s]);
    } else {
        e->rng -= r * icdf[s];
    }
    return enc_normalize(e);
}

/* encoder.rs:282-305 */
int orc_enc_bits(orc_enc *e, uint32_t fl, uint32_t bits)
{
    uint32_t window = e->end_window, used = e->end_bits;
    if (used + bits > WINDOW_SIZE) {
        do {
            if (enc_write_byte_at_end(e, (uint8_t)(window & SYM_MAX))) return e->error;
            window >>= SYM_BITS;
            used -= SYM_BITS;
        } while (used >= SYM_BITS);
    }
    window |= fl << used;
    used += bits;
    e->end_window = window;
    e->end_bits = used;
    e->bits_total += bits;
    return 0;
}

/* encoder.rs:258-274 */
int orc_enc_uint(orc_enc *e, uint32_t fl, uint32_t ft)
{
    ft -= 1u;
    uint32_t ftb = orc_ilog(ft);
    if (ftb > UINT_BITS) {
        ftb -= UINT_BITS;
        uint32_t ft1 = (ft >> ftb) + 1u;
        uint32_t fl1 = fl >> ftb;
        if (orc_enc_encode(e, fl1, fl1 + 1u, ft1)) return e->error;
        return orc_enc_bits(e, fl & ((1u << ftb) - 1u), ftb);
    }
    return orc_enc_encode(e, fl, fl + 1u, ft + 1u);
}

/* encoder.rs:327-347 */
int orc_enc_patch_initial_bits(orc_enc *e, uint32_t val, uint32_t nbits)
{
    uint32_t shift = SYM_BITS - nbits;
    uint32_t mask = ((1u << nbits) - 1u) << shift;
    if (e->offs > 0) {
        e->buffer[0] = (uint8_t)(((uint32_t)e->buffer[0] & ~mask) | (val << shift));
    } else if (e->rem >= 0) {
        e->rem = (int32_t)(((uint32_t)e->rem & ~mask) | (val << shift));
    } else if (e->rng <= (CODE_TOP >> nbits)) {
        e->val = (e->val & ~(mask << CODE_SHIFT)) | (val << (CODE_SHIFT + shift));
    } else {
        return e->error = ORC_ERR_INTERNAL;
    }
    return 0;
}

/* encoder.rs:361-369 */
void orc_enc_shrink(orc_enc *e, uint32_t len)
{
    uint32_t start = e->storage - e->end_offs;
    uint32_t dest = len - e->end_offs;
    memmove(e->buffer + dest, e->buffer + start, e->end_offs);
    e->storage = len;
}

/* encoder.rs:376-425 */
int orc_enc_done(orc_enc *e)
{
    int32_t l = (int32_t)(CODE_BITS - orc_ilog(e->rng));
    uint32_t mask = (CODE_TOP - 1u) >> l;
    uint32_t end = (e->val + mask) & ~mask;
    if ((end | mask) >= e->val + e->rng) {
        l += 1;
        mask >>= 1;
        end = (e->val + mask) & ~mask;
    }
    while (l > 0) {
        if (enc_carry_out(e, end >> CODE_SHIFT)) return e->error;
        end = (end << SYM_BITS) & (CODE_TOP - 1u);
        l -= (int32_t)SYM_BITS;
    }
    if (e->rem >= 0 || e->ext > 0) {
        if (enc_carry_out(e, 0)) return e->error;
    }
    uint32_t window = e->end_window, used = e->end_bits;
    while (used >= SYM_BITS) {
        if (enc_write_byte_at_end(e, (uint8_t)(window & SYM_MAX))) return e->error;
        window >>= SYM_BITS;
        used -= SYM_BITS;
    }
    memset(e->buffer + e->offs, 0, e->storage - e->end_offs - e->offs);
    if (used > 0) {
        if (e->end_offs >= e->storage) return e->error = ORC_ERR_INTERNAL;
        l = -l;
        if (e->offs + e->end_offs >= e->storage && l < (int32_t)used) window &= (1u << l) - 1u;
        e->buffer[e->storage - e->end_offs - 1u] |= (uint8_t)window;
    }
    return 0;
}

/* encoder.rs:437-482 */
int orc_enc_laplace(orc_enc *e, int32_t *value, uint32_t fs, uint32_t decay)
{
    int32_t val = *value;
    uint32_t fl = 0;
    if (val != 0) {
        int32_t s = val < 0 ? -1 : 0;
        val = (val + s) ^ s;
        fl = fs;
        fs = orc_laplace_freq1(fs, decay);
        int32_t i = 1;
        while (fs > 0 && i < val) {
            fs *= 2u;
            fl += fs + 2u;
            fs = (fs * decay) >> 15;
            i += 1;
        }
        if (fs == 0) {
            int32_t ndi_max = (int32_t)(32768u - fl);
            ndi_max = (ndi_max - s) >> 1;
            int32_t di = val - i < ndi_max - 1 ? val - i : ndi_max - 1;
            fl += (uint32_t)(2 * di + 1 + s);
            fs = umin(1u, 32768u - fl);
            *value = (i + di + s) ^ s;
        } else {
            fs += 1u;
            fl += (uint32_t)((int32_t)fs & ~s);
        }
    }
    return orc_enc_encode_bin(e, fl, fl + fs, 15);
}

uint32_t orc_enc_range_bytes(const orc_enc *e) { return e->offs; }
uint32_t orc_enc_tell(const orc_enc *e) { return orc_tell(e->bits_total, e->rng); }
uint32_t orc_enc_tell_frac(const orc_enc *e) { return orc_tell_frac(e->bits_total, e->rng); }

/* ------------------------------------------------------------------ scripts */

uint32_t orc_dec_run_script(const uint8_t *buf, uint32_t len, const orc_op *ops, uint32_t n_ops,
                            const uint8_t *icdf_pool, orc_op_out *out, int32_t *y_out)
{
    orc_dec d;
    uint32_t ny = 0;
    orc_dec_init(&d, buf, len);
    for (uint32_t i = 0; i < n_ops; i++) {
        uint32_t a = ops[i].a, b = ops[i].b, v = 0;
        switch (ops[i].op) {
        case ORC_OP_UINT: v = orc_dec_uint(&d, a); break;
        case ORC_OP_BITS: v = orc_dec_bits(&d, a); break;
        case ORC_OP_BIT_LOGP: v = (uint32_t)orc_dec_bit_logp(&d, a); break;
        case ORC_OP_ICDF: v = orc_dec_icdf(&d, icdf_pool + a, b); break;
        case ORC_OP_LAPLACE: v = (uint32_t)orc_dec_laplace(&d, a, b); break;
        case ORC_OP_BIT_VIA_DECODE: { /* mod.rs:446-454 */
            uint32_t fs = orc_dec_decode(&d, 1u << a);
            int s = fs >= (1u << a) - 1u;
            orc_dec_update(&d, s ? (1u << a) - 1u : 0u, (1u << a) - (s ? 0u : 1u), 1u << a);
            v = (uint32_t)s;
            break;
        }
        case ORC_OP_BIT_VIA_DECODE_BIN: { /* mod.rs:455-463 */
            uint32_t fs = orc_dec_decode_bin(&d, a);
            int s = fs >= (1u << a) - 1u;
            orc_dec_update(&d, s ? (1u << a) - 1u : 0u, (1u << a) - (s ? 0u : 1u), 1u << a);
            v = (uint32_t)s;
            break;
        }
        case ORC_OP_PULSES: {
            float yy = orc_decode_pulses(&d, y_out + ny, a, b);
            memcpy(&v, &yy, 4);
            ny += a;
            break;
        }
        case ORC_OP_SHRINK: orc_dec_shrink_storage(&d, a); break;
        case ORC_OP_TELL: v = orc_dec_tell(&d); break;
        default: break;
        }
        out[i].value = v;
        out[i].tell_frac = orc_dec_tell_frac(&d);
        out[i].rng = d.rng;
    }
    return ny;
}

int orc_enc_run_script(uint8_t *buf, uint32_t len, const orc_op *ops, const uint32_t *values,
                       uint32_t n_ops, const uint8_t *icdf_pool, const int32_t *y_in,
                       uint32_t *tell_frac_out, uint32_t *range_bytes, uint32_t *final_tell_frac)
{
    orc_enc e;
    uint32_t ny = 0;
    orc_enc_init(&e, buf, len);
    for (uint32_t i = 0; i < n_ops && !e.error; i++) {
        uint32_t a = ops[i].a, b = ops[i].b, v = values ? values[i] : 0;
        switch (ops[i].op) {
        case ORC_OP_UINT: orc_enc_uint(&e, v, a); break;
        case ORC_OP_BITS: orc_enc_bits(&e, v, a); break;
        case ORC_OP_BIT_LOGP: orc_enc_bit_logp(&e, v, a); break;
        case ORC_OP_ICDF: orc_enc_icdf(&e, v, icdf_pool + a, b); break;
        case ORC_OP_LAPLACE: { int32_t x = (int32_t)v; orc_enc_laplace(&e, &x, a, b); break; }
        case ORC_OP_BIT_VIA_DECODE: /* mod.rs:400-404 */
            orc_enc_encode(&e, v ? (1u << a) - 1u : 0u, (1u << a) - (v ? 0u : 1u), 1u << a);
            break;
        case ORC_OP_BIT_VIA_DECODE_BIN: /* mod.rs:405-409 */
            orc_enc_encode_bin(&e, v ? (1u << a) - 1u : 0u, (1u << a) - (v ? 0u : 1u), a);
            break;
        case ORC_OP_PULSES: orc_encode_pulses(&e, y_in + ny, a, b); ny += a; break;
        default: break;
        }
        if (tell_frac_out) tell_frac_out[i] = orc_enc_tell_frac(&e);
    }
    if (final_tell_frac) *final_tell_frac = orc_enc_tell_frac(&e);
    if (!e.error) orc_enc_done(&e);
    if (range_bytes) *range_bytes = e.offs;
    return e.error;
}

/* host_util.cpp -- logging, malloc-backed GString, Unicode predicates, UTF-8.
 *
 * The Unicode predicates restate what GLib's g_unichar_* functions compute from
 * the UCD (the reference calls them in chafa/chafa-symbol-map.c:477-500 and
 * chafa/internal/chafa-symbols.c:520-560).  Tables: unicode_ranges.inc, generated
 * by tools/gen_unicode_tables.py. */

#include "internal.h"

#include <cstdarg>
#include <mutex>

#include "unicode_ranges.inc"

/* ---- logging ---- */

bool
b200_verbose ()
{
    static int v = -1;
    if (v < 0)
    {
        const char *e = getenv ("CHAFA_B200_VERBOSE");
        v = (e && *e && *e != '0') ? 1 : 0;
    }
    return v == 1;
}

void
b200_log (const char *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    fputs ("chafa-b200: ", stderr);
    vfprintf (stderr, fmt, ap);
    fputc ('\n', stderr);
    va_end (ap);
}

static thread_local char last_error [512] = "";

void
b200_set_error (const char *fmt, ...)
{
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (last_error, sizeof (last_error), fmt, ap);
    va_end (ap);
    if (b200_verbose ())
        b200_log ("%s", last_error);
}

extern "C" const char *
chafa_b200_last_error (void)
{
    return last_error;
}

/* ---- memory / GString ---- */

void *
b200_malloc (size_t n)
{
    void *p = malloc (n ? n : 1);
    if (!p)
        abort ();
    return p;
}

void
b200_free (void *p)
{
    free (p);
}

/* GLib's g_string_free () ends up in free () for both the struct (g_slice falls
 * through to malloc in every GLib since 2.76, and is malloc-compatible with
 * G_SLICE=always-malloc before) and the character data, so malloc'd storage is
 * the compatible choice. */
GString *
b200_string_new_take (char *data, size_t len, size_t alloc)
{
    GString *gs = (GString *) b200_malloc (sizeof (GString));
    gs->str = data;
    gs->len = len;
    gs->allocated_len = alloc;
    return gs;
}

GString *
b200_string_new_copy (const char *data, size_t len)
{
    char *d = (char *) b200_malloc (len + 1);
    if (len)
        memcpy (d, data, len);
    d [len] = 0;
    return b200_string_new_take (d, len, len + 1);
}

extern "C" void
chafa_b200_string_free (GString *string)
{
    if (!string)
        return;
    free (string->str);
    free (string);
}

extern "C" void
chafa_b200_free (void *p)
{
    free (p);
}

static GQuark
quark_for (const char *domain)
{
    /* Stable small integers per domain string; only compared for equality by callers. */
    static std::map<std::string, GQuark> quarks;
    static std::mutex quarks_mutex;   /* g_quark_from_static_string is thread-safe; so is this */
    std::lock_guard<std::mutex> lock (quarks_mutex);
    auto it = quarks.find (domain);
    if (it != quarks.end ())
        return it->second;
    GQuark q = (GQuark) quarks.size () + 1;
    quarks [domain] = q;
    return q;
}

void
b200_set_gerror (GError **error, const char *domain, int code, const char *fmt, ...)
{
    if (!error)
        return;
    char buf [512];
    va_list ap;
    va_start (ap, fmt);
    vsnprintf (buf, sizeof (buf), fmt, ap);
    va_end (ap);
    GError *e = (GError *) b200_malloc (sizeof (GError));
    e->domain = quark_for (domain);
    e->code = code;
    e->message = strdup (buf);
    *error = e;
}

/* ---- unicode ---- */

static bool
in_ranges (uint32_t c, const unsigned int (*r) [2], int n)
{
    int lo = 0, hi = n - 1;
    while (lo <= hi)
    {
        int mid = (lo + hi) / 2;
        if (c < r [mid] [0])
            hi = mid - 1;
        else if (c > r [mid] [1])
            lo = mid + 1;
        else
            return true;
    }
    return false;
}

bool uc_isprint (uint32_t c) { return in_ranges (c, ucd_print_ranges, UCD_PRINT_N); }
bool uc_isalpha (uint32_t c) { return in_ranges (c, ucd_alpha_ranges, UCD_ALPHA_N); }
bool uc_isdigit (uint32_t c) { return in_ranges (c, ucd_digit_ranges, UCD_DIGIT_N); }
bool uc_ismark (uint32_t c) { return in_ranges (c, ucd_mark_ranges, UCD_MARK_N); }
bool uc_iswide (uint32_t c) { return in_ranges (c, ucd_wide_ranges, UCD_WIDE_N); }

bool
uc_iszerowidth (uint32_t c)
{
    if (c == 0x00ad)
        return false;
    if (in_ranges (c, ucd_zwtype_ranges, UCD_ZWTYPE_N))
        return true;
    return (c >= 0x1160 && c < 0x1200) || c == 0x200b;
}

bool
uc_iswide_cjk (uint32_t c)
{
    if (uc_iswide (c))
        return true;
    if (c == 0)
        return false;
    return in_ranges (c, ucd_ambiguous_ranges, UCD_AMBIGUOUS_N);
}

bool
uc_is_rtl_script (uint32_t c)
{
    return in_ranges (c, ucd_script_arabic_ranges, UCD_SCRIPT_ARABIC_N)
        || in_ranges (c, ucd_script_hebrew_ranges, UCD_SCRIPT_HEBREW_N)
        || in_ranges (c, ucd_script_syriac_ranges, UCD_SCRIPT_SYRIAC_N)
        || in_ranges (c, ucd_script_thaana_ranges, UCD_SCRIPT_THAANA_N);
}

int
uc_to_utf8 (uint32_t c, char *out)
{
    int len, first;

    if (c < 0x80) { first = 0; len = 1; }
    else if (c < 0x800) { first = 0xc0; len = 2; }
    else if (c < 0x10000) { first = 0xe0; len = 3; }
    else if (c < 0x200000) { first = 0xf0; len = 4; }
    else if (c < 0x4000000) { first = 0xf8; len = 5; }
    else { first = 0xfc; len = 6; }

    if (out)
    {
        for (int i = len - 1; i > 0; i--)
        {
            out [i] = (char) ((c & 0x3f) | 0x80);
            c >>= 6;
        }
        out [0] = (char) (c | first);
    }
    return len;
}

uint32_t
utf8_get_char (const char *p, int *len_out)
{
    const unsigned char *u = (const unsigned char *) p;
    uint32_t c = u [0];
    int len;

    if (c < 0x80) len = 1;
    else if ((c & 0xe0) == 0xc0) { len = 2; c &= 0x1f; }
    else if ((c & 0xf0) == 0xe0) { len = 3; c &= 0x0f; }
    else if ((c & 0xf8) == 0xf0) { len = 4; c &= 0x07; }
    else if ((c & 0xfc) == 0xf8) { len = 5; c &= 0x03; }
    else if ((c & 0xfe) == 0xfc) { len = 6; c &= 0x01; }
    else { if (len_out) *len_out = 1; return (uint32_t) -1; }

    for (int i = 1; i < len; i++)
    {
        if ((u [i] & 0xc0) != 0x80)
        {
            if (len_out) *len_out = i;
            return (uint32_t) -1;
        }
        c = (c << 6) | (u [i] & 0x3f);
    }
    if (len_out)
        *len_out = len;
    return c;
}

/* ------------------------------------------------------------------------ */

#include <chrono>
#include <map>

static const bool g_host_prof_on = getenv ("CHAFA_B200_HOST_PROFILE") != nullptr;
static std::mutex g_host_prof_mutex;
static std::map<std::string, std::pair<long long, long long>> *g_host_prof;

static void
host_prof_report ()
{
    std::lock_guard<std::mutex> lock (g_host_prof_mutex);
    if (!g_host_prof)
        return;
    for (const auto &e : *g_host_prof)
        fprintf (stderr, "chafa-b200 host profile: %-28s %9lld calls %10.3f us each\n", e.first.c_str (),
                 e.second.second, e.second.second > 30 ? 1e-3 * (double) e.second.first / (double) (e.second.second - 30) : 0.0);
}

static long long
host_prof_now ()
{
    return std::chrono::duration_cast<std::chrono::nanoseconds> (std::chrono::steady_clock::now ().time_since_epoch ()).count ();
}

HostProf::HostProf (const char *label_in) : label (label_in), t0 (g_host_prof_on ? host_prof_now () : 0)
{
}

HostProf::~HostProf ()
{
    if (!g_host_prof_on)
        return;
    long long dt = host_prof_now () - t0;
    std::lock_guard<std::mutex> lock (g_host_prof_mutex);
    if (!g_host_prof)
    {
        g_host_prof = new std::map<std::string, std::pair<long long, long long>> ();
        atexit (host_prof_report);
    }
    auto &e = (*g_host_prof) [label];
    e.second++;
    if (e.second > 30)          /* the first calls load modules, build tables and allocate */
        e.first += dt;
}

/* host_symbol_map.cpp -- ChafaSymbolMap: selectors, table build, host-side searches.
 *
 * Replaces chafa/chafa-symbol-map.c of the reference for the canvas path.  The
 * compiled table (code points + 8x8 bitmaps in EVALUATION ORDER) is what the
 * device kernels consume; building it is host work done once per canvas.
 *
 * Evaluation order matters for byte-identical output because every argmin in the
 * matcher breaks ties by position (chafa/internal/chafa-symbol-renderer.c:216).
 * The reference derives the order from a GLib GHashTable keyed by code point and
 * a qsort by popcount (chafa/chafa-symbol-map.c:384-470).  OrderedDirectSet below
 * reproduces GLib's published open-addressing scheme (prime modulus, quadratic
 * probing, in-place regrowth) so that the slot order -- and hence the table -- is
 * the same; the sort is a stable sort, which is what glibc's qsort does for
 * arrays of this size. */

#include "internal.h"

#include <algorithm>

#include "gen_tables.inc"

/* ------------------------------------------------------------------------ */
/* Slot-order emulation of a GHashTable with g_direct_hash keys               */

namespace {

class OrderedDirectSet
{
public:
    OrderedDirectSet () { set_shift (3); hashes.assign (size, 0); payload.assign (size, -1); keys.assign (size, 0); }

    /* Insert-or-replace; @value is an index owned by the caller. */
    void replace (uint32_t key, int value)
    {
        uint32_t h = key < 2 ? 2 : key;
        uint32_t idx = lookup (key, h);
        if (hashes [idx] >= 2)
        {
            payload [idx] = value;
            keys [idx] = key;
            return;
        }
        uint32_t old = hashes [idx];
        hashes [idx] = h;
        keys [idx] = key;
        payload [idx] = value;
        nnodes++;
        if (old == 0)
        {
            noccupied++;
            maybe_resize ();
        }
    }

    /* Values in slot order == g_hash_table_iter_next () order. */
    std::vector<int> values_in_order () const
    {
        std::vector<int> out;
        for (size_t i = 0; i < size; i++)
            if (hashes [i] >= 2)
                out.push_back (payload [i]);
        return out;
    }

private:
    size_t size = 0;
    uint32_t mod = 0, mask = 0;
    uint32_t nnodes = 0, noccupied = 0;
    std::vector<uint32_t> hashes, keys;
    std::vector<int> payload;

    void set_shift (int shift)
    {
        static const uint32_t prime_mod [] =
        {
            1, 2, 3, 7, 13, 31, 61, 127, 251, 509, 1021, 2039, 4093, 8191, 16381, 32749, 65521,
            131071, 262139, 524287, 1048573, 2097143, 4194301, 8388593, 16777213, 33554393,
            67108859, 134217689, 268435399, 536870909, 1073741789, 2147483647
        };
        size = (size_t) 1 << shift;
        mod = prime_mod [shift];
        mask = (uint32_t) size - 1;
    }

    uint32_t to_index (uint32_t h) const { return (h * 11u) % mod; }

    uint32_t lookup (uint32_t key, uint32_t h) const
    {
        uint32_t idx = to_index (h), step = 0;
        while (hashes [idx] != 0)
        {
            if (hashes [idx] == h && keys [idx] == key)
                return idx;
            step++;
            idx = (idx + step) & mask;
        }
        return idx;
    }

    void maybe_resize ()
    {
        if ((size > (size_t) nnodes * 4 && size > 8) || size <= noccupied + noccupied / 16)
            resize ();
    }

    void resize ()
    {
        size_t old_size = size;
        int n = (int) (nnodes * 1.333), shift = 0;
        while (n) { n >>= 1; shift++; }
        set_shift (std::max (shift, 3));

        if (size > old_size)
        {
            hashes.resize (size, 0);
            keys.resize (size, 0);
            payload.resize (size, -1);
        }
        std::vector<bool> moved (std::max (size, old_size), false);

        for (size_t i = 0; i < old_size; i++)
        {
            uint32_t h = hashes [i];
            if (h < 2) { hashes [i] = 0; continue; }
            if (moved [i]) continue;

            hashes [i] = 0;
            uint32_t k = keys [i];
            int v = payload [i];

            for (;;)
            {
                uint32_t idx = to_index (h), step = 0;
                while (moved [idx]) { step++; idx = (idx + step) & mask; }
                moved [idx] = true;
                uint32_t replaced = hashes [idx];
                hashes [idx] = h;
                if (replaced < 2)
                {
                    keys [idx] = k;
                    payload [idx] = v;
                    break;
                }
                std::swap (k, keys [idx]);
                std::swap (v, payload [idx]);
                h = replaced;
            }
        }
        if (size < old_size)
        {
            hashes.resize (size);
            keys.resize (size);
            payload.resize (size);
        }
        noccupied = nnodes;
    }
};

int
popcount64 (uint64_t v)
{
    return __builtin_popcountll (v);
}

} /* namespace */

/* ------------------------------------------------------------------------ */
/* Selection (reference: chafa/chafa-symbol-map.c:477-536)                    */

static bool
char_is_selected (const std::vector<Selector> &selectors, uint32_t tags, uint32_t c)
{
    uint32_t auto_exclude = CHAFA_SYMBOL_TAG_BAD;
    bool selected = false;

    if (!uc_isprint (c) || uc_iszerowidth (c) || c == '\t')
        return false;
    if (uc_is_rtl_script (c))
        return false;

    for (const Selector &s : selectors)
    {
        if (!s.is_range)
        {
            if (tags & s.tags)
            {
                selected = s.additive;
                auto_exclude &= ~s.tags;
            }
        }
        else if (c >= s.first && c <= s.last)
        {
            selected = s.additive;
        }
    }

    if (tags & auto_exclude)
        selected = false;
    return selected;
}

uint32_t
symbol_tags_for_char (uint32_t c)
{
    for (int i = 0; i < GEN_N_BUILTIN_NARROW; i++)
        if (gen_builtin_narrow [i].c == c)
            return gen_builtin_narrow [i].tags;
    for (int i = 0; i < GEN_N_BUILTIN_WIDE; i++)
        if (gen_builtin_wide [i].c == c)
            return gen_builtin_wide [i].tags;

    /* Default tags for a code point without a built-in glyph
     * (chafa/internal/chafa-symbols.c:520-560, the parts computable from the UCD). */
    uint32_t tags = 0;
    if (uc_iswide (c))
        tags |= CHAFA_SYMBOL_TAG_WIDE;
    else if (uc_iswide_cjk (c) && !((c >= 0xe000 && c <= 0xf8ff) || (c >= 0xf0000 && c <= 0x10ffff)))
        tags |= CHAFA_SYMBOL_TAG_AMBIGUOUS;
    if (uc_ismark (c) || uc_iszerowidth (c))
        tags |= CHAFA_SYMBOL_TAG_AMBIGUOUS;
    if (c <= 0x7f) tags |= CHAFA_SYMBOL_TAG_ASCII;
    else if (c >= 0x2300 && c <= 0x23ff) tags |= CHAFA_SYMBOL_TAG_TECHNICAL;
    else if (c >= 0x25a0 && c <= 0x25ff) tags |= CHAFA_SYMBOL_TAG_GEOMETRIC;
    else if (c >= 0x2800 && c <= 0x28ff) tags |= CHAFA_SYMBOL_TAG_BRAILLE;
    else if (c >= 0x1fb00 && c <= 0x1fb3b) tags |= CHAFA_SYMBOL_TAG_SEXTANT;
    if (uc_isalpha (c)) tags |= CHAFA_SYMBOL_TAG_ALPHA;
    if (uc_isdigit (c)) tags |= CHAFA_SYMBOL_TAG_DIGIT;
    if (!(tags & CHAFA_SYMBOL_TAG_WIDE))
        tags |= CHAFA_SYMBOL_TAG_NARROW;
    return tags;
}

/* ------------------------------------------------------------------------ */
/* Table build (reference: chafa/chafa-symbol-map.c:384-470, 560-678)         */

void
symbol_map_prepare (ChafaSymbolMap *map)
{
    if (!map->need_rebuild)
        return;

    std::vector<CompiledSymbol> pool;
    std::vector<CompiledSymbol2> pool2;
    OrderedDirectSet set, set2;

    if (map->use_builtin_glyphs)
    {
        for (int i = 0; i < GEN_N_BUILTIN_NARROW; i++)
        {
            if (!char_is_selected (map->selectors, gen_builtin_narrow [i].tags, gen_builtin_narrow [i].c))
                continue;
            CompiledSymbol s;
            s.c = gen_builtin_narrow [i].c;
            s.tags = gen_builtin_narrow [i].tags;
            s.bitmap = gen_builtin_narrow [i].bitmap;
            s.popcount = popcount64 (s.bitmap);
            pool.push_back (s);
            set.replace (s.c, (int) pool.size () - 1);
        }
        for (int i = 0; i < GEN_N_BUILTIN_WIDE; i++)
        {
            if (!char_is_selected (map->selectors, gen_builtin_wide [i].tags, gen_builtin_wide [i].c))
                continue;
            CompiledSymbol2 s;
            s.c = gen_builtin_wide [i].c;
            s.tags = gen_builtin_wide [i].tags;
            s.bitmap [0] = gen_builtin_wide [i].bitmap [0];
            s.bitmap [1] = gen_builtin_wide [i].bitmap [1];
            s.popcount = popcount64 (s.bitmap [0]) + popcount64 (s.bitmap [1]);
            pool2.push_back (s);
            set2.replace (s.c, (int) pool2.size () - 1);
        }
    }

    /* User glyphs: iterate in the glyph table's own slot order */
    {
        OrderedDirectSet gset;
        for (size_t i = 0; i < map->glyphs.size (); i++)
            gset.replace (map->glyphs [i].first, (int) i);
        for (int gi : gset.values_in_order ())
        {
            uint32_t c = map->glyphs [gi].first;
            uint32_t tags = symbol_tags_for_char (c) | CHAFA_SYMBOL_TAG_IMPORTED;
            if (!char_is_selected (map->selectors, tags, c))
                continue;
            CompiledSymbol s;
            s.c = c;
            s.tags = tags;
            s.bitmap = map->glyphs [gi].second;
            s.popcount = popcount64 (s.bitmap);
            pool.push_back (s);
            set.replace (c, (int) pool.size () - 1);
        }
    }
    {
        OrderedDirectSet gset;
        for (size_t i = 0; i < map->glyphs2.size (); i++)
            gset.replace (map->glyphs2 [i].first, (int) i);
        for (int gi : gset.values_in_order ())
        {
            uint32_t c = map->glyphs2 [gi].first;
            uint32_t tags = symbol_tags_for_char (c) | CHAFA_SYMBOL_TAG_IMPORTED;
            if (!char_is_selected (map->selectors, tags, c))
                continue;
            CompiledSymbol2 s;
            s.c = c;
            s.tags = tags;
            s.bitmap [0] = map->glyphs2 [gi].second.first;
            s.bitmap [1] = map->glyphs2 [gi].second.second;
            s.popcount = popcount64 (s.bitmap [0]) + popcount64 (s.bitmap [1]);
            pool2.push_back (s);
            set2.replace (c, (int) pool2.size () - 1);
        }
    }

    map->symbols.clear ();
    for (int i : set.values_in_order ())
        map->symbols.push_back (pool [i]);
    std::stable_sort (map->symbols.begin (), map->symbols.end (),
                      [] (const CompiledSymbol &a, const CompiledSymbol &b) { return a.popcount < b.popcount; });

    map->symbols2.clear ();
    for (int i : set2.values_in_order ())
        map->symbols2.push_back (pool2 [i]);
    std::stable_sort (map->symbols2.begin (), map->symbols2.end (),
                      [] (const CompiledSymbol2 &a, const CompiledSymbol2 &b) { return a.popcount < b.popcount; });

    map->need_rebuild = false;
}

bool
symbol_map_has_symbol (const ChafaSymbolMap *map, uint32_t c)
{
    for (const CompiledSymbol &s : map->symbols)
        if (s.c == c)
            return true;
    for (const CompiledSymbol2 &s : map->symbols2)
        if (s.c == c)
            return true;
    return false;
}

/* ------------------------------------------------------------------------ */
/* Host searches used at canvas creation (blank/solid char): same ordering     */
/* rules as the device kernels (chafa/chafa-symbol-map.c:1153-1253, 1335-1426) */

int
symbol_map_find_candidates (const ChafaSymbolMap *map, uint64_t bitmap, bool do_inverse,
                            Candidate *out, int n_max)
{
    /* Stable top-8 over the sequence (i, plain), (i, inverted), ... keyed by distance */
    std::vector<std::pair<int, int>> keyed;  /* (hd, seq) */
    for (size_t i = 0; i < map->symbols.size (); i++)
    {
        int hd = popcount64 (bitmap ^ map->symbols [i].bitmap);
        keyed.push_back ({ hd, (int) i * 2 });
        if (do_inverse)
            keyed.push_back ({ 64 - hd, (int) i * 2 + 1 });
    }
    std::stable_sort (keyed.begin (), keyed.end (),
                      [] (const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
    int n = std::min ({ (int) keyed.size (), N_CANDIDATES_MAX, n_max });
    for (int i = 0; i < n; i++)
    {
        out [i].symbol_index = keyed [i].second >> 1;
        out [i].hamming_distance = keyed [i].first;
        out [i].is_inverted = keyed [i].second & 1;
    }
    return n;
}

static int
find_closest_popcount (const ChafaSymbolMap *map, int popcount)
{
    int n = (int) map->symbols.size ();
    int i = 0, j = n - 1;

    while (i < j)
    {
        int k = (i + j + 1) / 2;
        if (popcount < map->symbols [k].popcount)
            j = k - 1;
        else
            i = k;
    }
    if (i < n - 1
        && abs (popcount - map->symbols [i + 1].popcount) < abs (popcount - map->symbols [i].popcount))
        i++;
    return i;
}

int
symbol_map_find_fill_candidates (const ChafaSymbolMap *map, int popcount, bool do_inverse,
                                 Candidate *out, int n_max)
{
    if (n_max <= 0 || map->symbols.empty ())
        return 0;

    int sym = find_closest_popcount (map, popcount);
    out [0].symbol_index = sym;
    out [0].hamming_distance = abs (popcount - map->symbols [sym].popcount);
    out [0].is_inverted = false;

    if (do_inverse && out [0].hamming_distance != 0)
    {
        sym = find_closest_popcount (map, 64 - popcount);
        int d = abs (64 - popcount - map->symbols [sym].popcount);
        if (d < out [0].hamming_distance)
        {
            out [0].symbol_index = sym;
            out [0].hamming_distance = d;
            out [0].is_inverted = true;
        }
    }
    return 1;
}

/* ------------------------------------------------------------------------ */
/* Selector string parsing (reference: chafa/chafa-symbol-map.c:800-1043)     */

static bool
parse_code_point (const char *str, int len, int *parsed_len_out, uint32_t *c_out)
{
    int i = 0;
    uint32_t code = 0;
    bool result = false;

    if (len >= 1 && (str [0] == 'u' || str [0] == 'U'))
        i++;
    if (len >= 2 && str [0] == '0' && str [1] == 'x')
        i += 2;

    for ( ; i < len; i++)
    {
        int c = (unsigned char) str [i];
        if (c >= '0' && c <= '9') code = code * 16 + (c - '0');
        else if (c >= 'a' && c <= 'f') code = code * 16 + (c - 'a' + 10);
        else if (c >= 'A' && c <= 'F') code = code * 16 + (c - 'A' + 10);
        else break;
        result = true;
    }

    *parsed_len_out = i;
    *c_out = code;
    return result;
}

static bool
parse_symbol_tag (const char *name, int len, bool *is_range_out, uint32_t *tags_out,
                  uint32_t *first_out, uint32_t *last_out, GError **error)
{
    static const struct { const char *name; uint32_t tags; } map [] =
    {
        { "all", (uint32_t) CHAFA_SYMBOL_TAG_ALL }, { "none", CHAFA_SYMBOL_TAG_NONE },
        { "space", CHAFA_SYMBOL_TAG_SPACE }, { "solid", CHAFA_SYMBOL_TAG_SOLID },
        { "stipple", CHAFA_SYMBOL_TAG_STIPPLE }, { "block", CHAFA_SYMBOL_TAG_BLOCK },
        { "border", CHAFA_SYMBOL_TAG_BORDER }, { "diagonal", CHAFA_SYMBOL_TAG_DIAGONAL },
        { "dot", CHAFA_SYMBOL_TAG_DOT }, { "quad", CHAFA_SYMBOL_TAG_QUAD },
        { "half", CHAFA_SYMBOL_TAG_HALF }, { "hhalf", CHAFA_SYMBOL_TAG_HHALF },
        { "vhalf", CHAFA_SYMBOL_TAG_VHALF }, { "inverted", CHAFA_SYMBOL_TAG_INVERTED },
        { "braille", CHAFA_SYMBOL_TAG_BRAILLE }, { "sextant", CHAFA_SYMBOL_TAG_SEXTANT },
        { "wedge", CHAFA_SYMBOL_TAG_WEDGE }, { "technical", CHAFA_SYMBOL_TAG_TECHNICAL },
        { "geometric", CHAFA_SYMBOL_TAG_GEOMETRIC }, { "ascii", CHAFA_SYMBOL_TAG_ASCII },
        { "alpha", CHAFA_SYMBOL_TAG_ALPHA }, { "digit", CHAFA_SYMBOL_TAG_DIGIT },
        { "narrow", CHAFA_SYMBOL_TAG_NARROW }, { "wide", CHAFA_SYMBOL_TAG_WIDE },
        { "ambiguous", CHAFA_SYMBOL_TAG_AMBIGUOUS }, { "ugly", CHAFA_SYMBOL_TAG_UGLY },
        { "extra", CHAFA_SYMBOL_TAG_EXTRA }, { "alnum", CHAFA_SYMBOL_TAG_ALNUM },
        { "bad", CHAFA_SYMBOL_TAG_BAD }, { "legacy", CHAFA_SYMBOL_TAG_LEGACY },
        { "latin", CHAFA_SYMBOL_TAG_LATIN }, { "import", CHAFA_SYMBOL_TAG_IMPORTED },
        { "imported", CHAFA_SYMBOL_TAG_IMPORTED }, { "octant", CHAFA_SYMBOL_TAG_OCTANT },
        { nullptr, 0 }
    };

    /* A tag matches when the given text is a case-insensitive prefix of its name;
     * the first table entry that matches wins. */
    for (int i = 0; map [i].name; i++)
    {
        if (!strncasecmp (map [i].name, name, len))
        {
            *tags_out = map [i].tags;
            *is_range_out = false;
            return true;
        }
    }

    int parsed_len;
    if (!parse_code_point (name, len, &parsed_len, first_out))
        goto err;

    if (len - parsed_len > 0)
    {
        int parsed_last_len;
        if (len - parsed_len < 3
            || name [parsed_len] != '.' || name [parsed_len + 1] != '.'
            || !parse_code_point (name + parsed_len + 2, len - parsed_len - 2, &parsed_last_len, last_out)
            || parsed_len + 2 + parsed_last_len != len)
            goto err;
    }
    else
    {
        *last_out = *first_out;
    }

    *is_range_out = true;
    return true;

err:
    b200_set_gerror (error, "g-option-context-error-quark", 1, "Unrecognized symbol tag '%.*s'.", len, name);
    return false;
}

static bool
parse_selectors (ChafaSymbolMap *map, const char *selectors, GError **error)
{
    const char *p0 = selectors;
    bool is_add = false, is_remove = false, do_clear = false;
    std::vector<Selector> parsed;

    auto push = [&] (bool is_range, uint32_t tags, uint32_t first, uint32_t last)
    {
        Selector s;
        s.is_range = is_range;
        s.additive = is_add;
        s.tags = tags;
        s.first = first;
        s.last = last;
        parsed.push_back (s);
    };

    while (*p0)
    {
        int n;

        p0 += strspn (p0, " ,");
        if (!*p0)
            break;

        if (*p0 == '-') { is_add = false; is_remove = true; p0++; }
        else if (*p0 == '+') { is_add = true; is_remove = false; p0++; }

        p0 += strspn (p0, " ");
        if (!*p0)
            break;

        if (!is_add && !is_remove)
        {
            do_clear = true;
            is_add = true;
        }

        if (*p0 == '[')
        {
            bool escape = false;
            p0++;

            while (*p0)
            {
                int clen;
                uint32_t c = utf8_get_char (p0, &clen);
                /* Advance by the lead byte's nominal length, as g_utf8_next_char () does */
                unsigned char lead = (unsigned char) *p0;
                int skip = lead < 0xc0 ? 1 : lead < 0xe0 ? 2 : lead < 0xf0 ? 3 : lead < 0xf8 ? 4 : lead < 0xfc ? 5 : 6;

                if (c == '\\' && !escape)
                {
                    escape = true;
                    p0 += skip;
                    continue;
                }
                if (c == ']' && !escape)
                    break;

                push (true, 0, c, c);
                escape = false;

                /* Do not run past the terminator on a truncated sequence */
                for (int k = 0; k < skip && *p0; k++)
                    p0++;
            }

            if (*p0 != ']')
            {
                b200_set_gerror (error, "g-option-context-error-quark", 1,
                                 "Syntax error in symbol selector set.");
                return false;
            }
            n = 1;
        }
        else if (!(n = (int) strspn (p0, "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789.")))
        {
            b200_set_gerror (error, "g-option-context-error-quark", 1,
                             "Syntax error in symbol tag selectors.");
            return false;
        }
        else
        {
            bool is_range;
            uint32_t tags = 0, first = 0, last = 0;
            if (!parse_symbol_tag (p0, n, &is_range, &tags, &first, &last, error))
                return false;
            push (is_range, tags, first, last);
        }

        p0 += n;
    }

    if (do_clear)
        map->selectors.clear ();
    map->selectors.insert (map->selectors.end (), parsed.begin (), parsed.end ());
    map->need_rebuild = true;
    return true;
}

/* ------------------------------------------------------------------------ */
/* Public API (reference: chafa/chafa-symbol-map.c:1539-1930)                 */

extern "C" {

ChafaSymbolMap *
chafa_symbol_map_new (void)
{
    return new ChafaSymbolMap ();
}

ChafaSymbolMap *
chafa_symbol_map_copy (const ChafaSymbolMap *symbol_map)
{
    RETURN_VAL_IF_FAIL (symbol_map != NULL, NULL);
    ChafaSymbolMap *m = new ChafaSymbolMap (*symbol_map);
    if (!symbol_map->need_rebuild)
        symbol_map_prepare (m);
    return m;
}

void
chafa_symbol_map_ref (ChafaSymbolMap *symbol_map)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    RETURN_IF_FAIL (symbol_map->refs.load () > 0);
    symbol_map->refs++;
}

void
chafa_symbol_map_unref (ChafaSymbolMap *symbol_map)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    RETURN_IF_FAIL (symbol_map->refs.load () > 0);
    if (--symbol_map->refs == 0)
        delete symbol_map;
}

static void
push_selector (ChafaSymbolMap *map, bool is_range, bool additive, uint32_t tags, uint32_t first, uint32_t last)
{
    Selector s;
    s.is_range = is_range;
    s.additive = additive;
    s.tags = tags;
    s.first = first;
    s.last = last;
    map->selectors.push_back (s);
    map->need_rebuild = true;
}

void
chafa_symbol_map_add_by_tags (ChafaSymbolMap *symbol_map, ChafaSymbolTags tags)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    RETURN_IF_FAIL (symbol_map->refs.load () > 0);
    push_selector (symbol_map, false, true, (uint32_t) tags, 0, 0);
}

void
chafa_symbol_map_remove_by_tags (ChafaSymbolMap *symbol_map, ChafaSymbolTags tags)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    RETURN_IF_FAIL (symbol_map->refs.load () > 0);
    push_selector (symbol_map, false, false, (uint32_t) tags, 0, 0);
}

void
chafa_symbol_map_add_by_range (ChafaSymbolMap *symbol_map, gunichar first, gunichar last)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    RETURN_IF_FAIL (symbol_map->refs.load () > 0);
    push_selector (symbol_map, true, true, 0, first, last);
}

void
chafa_symbol_map_remove_by_range (ChafaSymbolMap *symbol_map, gunichar first, gunichar last)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    RETURN_IF_FAIL (symbol_map->refs.load () > 0);
    push_selector (symbol_map, true, false, 0, first, last);
}

gboolean
chafa_symbol_map_apply_selectors (ChafaSymbolMap *symbol_map, const gchar *selectors, GError **error)
{
    RETURN_VAL_IF_FAIL (symbol_map != NULL, 0);
    RETURN_VAL_IF_FAIL (symbol_map->refs.load () > 0, 0);
    RETURN_VAL_IF_FAIL (selectors != NULL, 0);
    return parse_selectors (symbol_map, selectors, error) ? 1 : 0;
}

gboolean
chafa_symbol_map_get_allow_builtin_glyphs (ChafaSymbolMap *symbol_map)
{
    RETURN_VAL_IF_FAIL (symbol_map != NULL, 0);
    return symbol_map->use_builtin_glyphs ? 1 : 0;
}

void
chafa_symbol_map_set_allow_builtin_glyphs (ChafaSymbolMap *symbol_map, gboolean use_builtin_glyphs)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    if (symbol_map->use_builtin_glyphs == !!use_builtin_glyphs)
        return;
    symbol_map->use_builtin_glyphs = !!use_builtin_glyphs;
    symbol_map->need_rebuild = true;
}

/* User glyphs (reference: chafa/chafa-symbol-map.c:1794-1930).  The import itself --
 * scale to 8x8 / 16x8, 3x3 sharpen, threshold -- runs on the device (host_glyph.cpp). */
void
chafa_symbol_map_add_glyph (ChafaSymbolMap *symbol_map, gunichar code_point, ChafaPixelType pixel_format,
                            gpointer pixels, gint width, gint height, gint rowstride)
{
    RETURN_IF_FAIL (symbol_map != NULL);
    RETURN_IF_FAIL (pixels != NULL && width > 0 && height > 0);

    bool wide = uc_iswide (code_point);
    uint64_t bm [2] = { 0, 0 };

    if (!glyph_import_device (pixel_format, pixels, width, height, rowstride, wide, bm))
        return;

    if (wide)
    {
        bool found = false;
        for (auto &g : symbol_map->glyphs2)
            if (g.first == code_point)
            {
                g.second = { bm [0], bm [1] };
                found = true;
            }
        if (!found)
            symbol_map->glyphs2.push_back ({ code_point, { bm [0], bm [1] } });
    }
    else
    {
        bool found = false;
        for (auto &g : symbol_map->glyphs)
            if (g.first == code_point)
            {
                g.second = bm [0];
                found = true;
            }
        if (!found)
            symbol_map->glyphs.push_back ({ code_point, bm [0] });
    }
    symbol_map->need_rebuild = true;
}

gboolean
chafa_symbol_map_get_glyph (ChafaSymbolMap *symbol_map, gunichar code_point, ChafaPixelType pixel_format,
                            gpointer *pixels_out, gint *width_out, gint *height_out, gint *rowstride_out)
{
    RETURN_VAL_IF_FAIL (symbol_map != NULL, 0);

    uint64_t bm [2] = { 0, 0 };
    int n_halves = 0;

    if (uc_iswide (code_point))
    {
        for (const auto &g : symbol_map->glyphs2)
            if (g.first == code_point)
            {
                bm [0] = g.second.first;
                bm [1] = g.second.second;
                n_halves = 2;
            }
    }
    else
    {
        for (const auto &g : symbol_map->glyphs)
            if (g.first == code_point)
            {
                bm [0] = g.second;
                n_halves = 1;
            }
    }
    if (!n_halves)
        return 0;

    int width = 8 * n_halves, rowstride = width * 4;

    if (pixels_out)
    {
        /* A set pixel is opaque white, a clear one transparent black: the same bytes
         * in every 32-bit layout; 24-bit layouts keep the 32-bit rowstride. */
        int bpp = (pixel_format == CHAFA_PIXEL_RGB8 || pixel_format == CHAFA_PIXEL_BGR8) ? 3 : 4;
        uint8_t *px = (uint8_t *) b200_malloc ((size_t) rowstride * 8);
        memset (px, 0, (size_t) rowstride * 8);
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < width; x++)
            {
                uint64_t half = bm [x / 8];
                bool set = (half >> (63 - (y * 8 + (x & 7)))) & 1;
                if (set)
                    memset (px + (size_t) y * rowstride + (size_t) x * bpp, 0xff, bpp);
            }
        *pixels_out = px;
    }
    if (width_out) *width_out = width;
    if (height_out) *height_out = 8;
    if (rowstride_out) *rowstride_out = rowstride;
    return 1;
}

int
chafa_b200_symbol_map_get_compiled (const ChafaSymbolMap *map, int wide, int max_out,
                                    uint32_t *c_out, uint64_t *bitmap0_out, uint64_t *bitmap1_out)
{
    RETURN_VAL_IF_FAIL (map != NULL, 0);
    ChafaSymbolMap tmp (*map);
    symbol_map_prepare (&tmp);

    int n = wide ? (int) tmp.symbols2.size () : (int) tmp.symbols.size ();
    for (int i = 0; i < n && i < max_out; i++)
    {
        if (!wide)
        {
            c_out [i] = tmp.symbols [i].c;
            bitmap0_out [i] = tmp.symbols [i].bitmap;
            bitmap1_out [i] = 0;
        }
        else
        {
            c_out [i] = tmp.symbols2 [i].c;
            bitmap0_out [i] = tmp.symbols2 [i].bitmap [0];
            bitmap1_out [i] = tmp.symbols2 [i].bitmap [1];
        }
    }
    return n;
}

} /* extern "C" */

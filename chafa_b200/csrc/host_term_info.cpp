/* host_term_info.cpp -- ChafaTermInfo (emit half) and the built-in fallback description.
 *
 * Replaces the parts of chafa/chafa-term-info.c (template parsing 304-411,
 * emitters 413-445, set/get 1180-1315, generic emit 1336-1440) and
 * chafa/chafa-term-db.c (fallback info 569-582, 1316-1327) that the print path
 * uses.  Sequence templates use %1..%7 for arguments and %% for a literal '%';
 * parsing produces a literal byte string plus, per argument slot, the number of
 * literal bytes preceding it -- the same compiled form the device printer uses. */

#include "internal.h"

#include <cstdarg>

#include "gen_tables.inc"

#define ARG_INDEX_SENTINEL 255

static int
meta_n_args (int seq)
{
    return gen_term_seq_meta [seq].n_args;
}

static int
meta_type_size (int seq)
{
    const char *t = gen_term_seq_meta [seq].arg_type;
    if (!strcmp (t, "guint8")) return 1;
    if (!strcmp (t, "guint16")) return 2;
    if (!strcmp (t, "char")) return 1;
    return 4;
}

static bool
parse_seq (ParsedSeq *out, const char *in, int n_args, int arg_len_max, GError **error)
{
    char str [CHAFA_TERM_SEQ_LENGTH_MAX];
    SeqArg args [CHAFA_TERM_SEQ_ARGS_MAX];
    int i, j = 0, k = 0, pre_len = 0;

    if (n_args < 0)
        n_args = CHAFA_TERM_SEQ_ARGS_MAX - 1;

    memset (str, 0, sizeof (str));
    for (int a = 0; a < CHAFA_TERM_SEQ_ARGS_MAX; a++)
    {
        args [a].is_varargs = false;
        args [a].pre_len = 0;
        args [a].arg_index = ARG_INDEX_SENTINEL;
    }

    for (i = 0; j < CHAFA_TERM_SEQ_LENGTH_MAX && k < CHAFA_TERM_SEQ_ARGS_MAX && in [i]; i++)
    {
        char c = in [i];

        if (c == '%')
        {
            i++;
            c = in [i];

            if (c == '%')
            {
                str [j++] = '%';
                pre_len++;
            }
            else if (c >= '1' && c <= '7')
            {
                args [k].arg_index = c - '1';
                args [k].pre_len = pre_len;
                if (args [k].arg_index >= n_args)
                {
                    b200_set_gerror (error, "chafa-term-info-error-quark", CHAFA_TERM_INFO_ERROR_BAD_ARGUMENTS,
                                     "Control sequence had too many arguments.");
                    return false;
                }
                pre_len = 0;
                k++;
            }
            else if (c == 'v')
            {
                args [k].is_varargs = true;
                args [k].arg_index = 0;
                args [k].pre_len = pre_len;
                pre_len = 0;
                k++;
            }
            else
            {
                /* Bad escape: the reference fails here without setting an error */
                return false;
            }
        }
        else
        {
            str [j++] = c;
            pre_len++;
        }
    }

    if (k == CHAFA_TERM_SEQ_ARGS_MAX)
    {
        b200_set_gerror (error, "chafa-term-info-error-quark", CHAFA_TERM_INFO_ERROR_BAD_ARGUMENTS,
                         "Control sequence had too many arguments.");
        return false;
    }

    if (j + k * arg_len_max + 1 > CHAFA_TERM_SEQ_LENGTH_MAX)
    {
        b200_set_gerror (error, "chafa-term-info-error-quark", CHAFA_TERM_INFO_ERROR_SEQ_TOO_LONG,
                         "Control sequence too long.");
        return false;
    }

    args [k].pre_len = pre_len;
    args [k].arg_index = ARG_INDEX_SENTINEL;

    memcpy (out->str, str, sizeof (str));
    memcpy (out->args, args, sizeof (args));
    out->defined = true;
    out->unparsed = in;
    return true;
}

static void
clear_seq (ParsedSeq *s)
{
    s->defined = false;
    s->unparsed.clear ();
    memset (s->str, 0, sizeof (s->str));
    for (int a = 0; a < CHAFA_TERM_SEQ_ARGS_MAX; a++)
    {
        s->args [a].is_varargs = false;
        s->args [a].pre_len = 0;
        s->args [a].arg_index = ARG_INDEX_SENTINEL;
    }
}

static char *
format_dec_0_to_9999 (char *dest, unsigned arg)
{
    unsigned m = arg < 9999 ? arg : 9999;
    char tmp [4];
    int n = 0;
    do { tmp [n++] = (char) ('0' + m % 10); m /= 10; } while (m);
    while (n) *dest++ = tmp [--n];
    return dest;
}

static char *
format_dec_u8 (char *dest, uint8_t v)
{
    if (v >= 100) *dest++ = (char) ('0' + v / 100);
    if (v >= 10) *dest++ = (char) ('0' + (v / 10) % 10);
    *dest++ = (char) ('0' + v % 10);
    return dest;
}

static char *
format_hex_u16 (char *dest, uint16_t v)
{
    static const char hex [] = "0123456789abcdef";
    *dest++ = hex [(v >> 12) & 0xf];
    *dest++ = hex [(v >> 8) & 0xf];
    *dest++ = hex [(v >> 4) & 0xf];
    *dest++ = hex [v & 0xf];
    return dest;
}

/* Host-side formatter used by chafa_term_info_emit_seq () and by the sixel header
 * (which is a few dozen bytes and is assembled on the host). */
char *
term_info_emit (const ChafaTermInfo *ti, char *out, int seq, const unsigned *argv, int n_args, int type_size)
{
    const ParsedSeq *s = &ti->seqs [seq];
    int ofs = 0, i;

    if (n_args == 0)
    {
        memcpy (out, s->str, s->args [0].pre_len);
        return out + s->args [0].pre_len;
    }
    if (s->args [0].arg_index == ARG_INDEX_SENTINEL)
        return out;

    for (i = 0; i < n_args; i++)
    {
        memcpy (out, s->str + ofs, s->args [i].pre_len);
        out += s->args [i].pre_len;
        ofs += s->args [i].pre_len;
        unsigned v = argv [s->args [i].arg_index];
        if (type_size == 1) out = format_dec_u8 (out, (uint8_t) v);
        else if (type_size == 2) out = format_hex_u16 (out, (uint16_t) v);
        else out = format_dec_0_to_9999 (out, v);
    }
    memcpy (out, s->str + ofs, s->args [i].pre_len);
    return out + s->args [i].pre_len;
}

/* A sequence with a variable number of arguments ("%v": the arguments from there on, separated
 * by ';'; chafa/chafa-term-info.c:448-497) */
static char *
term_info_emit_varargs (const ChafaTermInfo *ti, char *out, int seq, const unsigned *argv, int n_args)
{
    const ParsedSeq *s = &ti->seqs [seq];
    bool in_list = false;
    int ofs = 0, i = 0, j = 0;

    if (s->args [0].arg_index == ARG_INDEX_SENTINEL)
        return out;

    while (i + j < n_args)
    {
        if (in_list)
        {
            *out++ = ';';
        }
        else
        {
            in_list = s->args [i].is_varargs;
            memcpy (out, s->str + ofs, s->args [i].pre_len);
            out += s->args [i].pre_len;
            ofs += s->args [i].pre_len;
        }
        out = format_dec_0_to_9999 (out, argv [s->args [i].arg_index + j]);
        if (in_list)
            j++;
        else
            i++;
    }
    if (in_list)
        i++;
    memcpy (out, s->str + ofs, s->args [i].pre_len);
    return out + s->args [i].pre_len;
}

static ChafaTermInfo *
build_fallback ()
{
    ChafaTermInfo *ti = new ChafaTermInfo ();
    for (int i = 0; i < CHAFA_TERM_SEQ_MAX; i++)
    {
        clear_seq (&ti->seqs [i]);
        if (gen_fallback_seqs [i])
            parse_seq (&ti->seqs [i], gen_fallback_seqs [i], meta_n_args (i),
                       meta_type_size (i) == 1 ? 3 : 4, NULL);
    }
    return ti;
}

extern "C" {

GQuark
chafa_term_info_error_quark (void)
{
    GError *e = NULL;
    b200_set_gerror (&e, "chafa-term-info-error-quark", 0, "x");
    GQuark q = e->domain;
    free (e->message);
    free (e);
    return q;
}

ChafaTermInfo *
chafa_term_info_new (void)
{
    ChafaTermInfo *ti = new ChafaTermInfo ();
    for (int i = 0; i < CHAFA_TERM_SEQ_MAX; i++)
        clear_seq (&ti->seqs [i]);
    return ti;
}

ChafaTermInfo *
chafa_term_info_copy (const ChafaTermInfo *term_info)
{
    RETURN_VAL_IF_FAIL (term_info != NULL, NULL);
    ChafaTermInfo *ti = new ChafaTermInfo ();
    for (int i = 0; i < CHAFA_TERM_SEQ_MAX; i++)
        ti->seqs [i] = term_info->seqs [i];
    return ti;
}

void
chafa_term_info_ref (ChafaTermInfo *term_info)
{
    RETURN_IF_FAIL (term_info != NULL);
    RETURN_IF_FAIL (term_info->refs.load () > 0);
    term_info->refs++;
}

void
chafa_term_info_unref (ChafaTermInfo *term_info)
{
    RETURN_IF_FAIL (term_info != NULL);
    RETURN_IF_FAIL (term_info->refs.load () > 0);
    if (--term_info->refs == 0)
        delete term_info;
}

const gchar *
chafa_term_info_get_seq (ChafaTermInfo *term_info, ChafaTermSeq seq)
{
    RETURN_VAL_IF_FAIL (term_info != NULL, NULL);
    RETURN_VAL_IF_FAIL ((int) seq >= 0 && seq < CHAFA_TERM_SEQ_MAX, NULL);
    return term_info->seqs [seq].defined ? term_info->seqs [seq].unparsed.c_str () : NULL;
}

gint
chafa_term_info_set_seq (ChafaTermInfo *term_info, ChafaTermSeq seq, const gchar *str, GError **error)
{
    RETURN_VAL_IF_FAIL (term_info != NULL, 0);
    RETURN_VAL_IF_FAIL ((int) seq >= 0 && seq < CHAFA_TERM_SEQ_MAX, 0);

    if (!str)
    {
        clear_seq (&term_info->seqs [seq]);
        return 1;
    }

    ParsedSeq tmp;
    clear_seq (&tmp);
    if (!parse_seq (&tmp, str, meta_n_args (seq), meta_type_size (seq) == 1 ? 3 : 4, error))
        return 0;
    term_info->seqs [seq] = tmp;
    return 1;
}

gboolean
chafa_term_info_have_seq (const ChafaTermInfo *term_info, ChafaTermSeq seq)
{
    RETURN_VAL_IF_FAIL (term_info != NULL, 0);
    RETURN_VAL_IF_FAIL ((int) seq >= 0 && seq < CHAFA_TERM_SEQ_MAX, 0);
    return term_info->seqs [seq].defined ? 1 : 0;
}

void
chafa_term_info_supplement (ChafaTermInfo *term_info, ChafaTermInfo *source)
{
    RETURN_IF_FAIL (term_info != NULL);
    RETURN_IF_FAIL (source != NULL);
    for (int i = 0; i < CHAFA_TERM_SEQ_MAX; i++)
        if (!term_info->seqs [i].defined && source->seqs [i].defined)
            term_info->seqs [i] = source->seqs [i];
}

gchar *
chafa_term_info_emit_seq (ChafaTermInfo *term_info, ChafaTermSeq seq, ...)
{
    RETURN_VAL_IF_FAIL (term_info != NULL, NULL);
    RETURN_VAL_IF_FAIL ((int) seq >= 0 && seq < CHAFA_TERM_SEQ_MAX, NULL);

    unsigned argv [CHAFA_TERM_SEQ_ARGS_MAX];
    int n = 0, want = meta_n_args (seq), tsz = meta_type_size (seq);
    bool bad = false;
    va_list ap;

    va_start (ap, seq);
    for (;;)
    {
        int arg = va_arg (ap, int);
        if (arg < 0)
            break;
        if (n == CHAFA_TERM_SEQ_ARGS_MAX || (want >= 0 && n == want)) { bad = true; break; }
        if ((tsz == 1 && arg > 0xff) || (tsz == 2 && arg > 0xffff)) { bad = true; break; }
        argv [n++] = (unsigned) arg;
    }
    va_end (ap);

    if (bad || (want >= 0 && n != want) || want < 0 /* varargs sequences are not on the print path */)
        return NULL;
    if (!term_info->seqs [seq].defined)
        return NULL;

    char buf [CHAFA_TERM_SEQ_LENGTH_MAX + 8];
    char *p = term_info_emit (term_info, buf, seq, argv, n, tsz);
    size_t len = (size_t) (p - buf);
    char *res = (char *) b200_malloc (len + 1);
    memcpy (res, buf, len);
    res [len] = 0;
    return res;
}

#include "term_seq_emit.inc"

ChafaTermDb *
chafa_term_db_get_default (void)
{
    static ChafaTermDb db = { nullptr };
    static std::atomic<bool> ready { false };
    static std::atomic_flag lock = ATOMIC_FLAG_INIT;

    if (!ready.load ())
    {
        while (lock.test_and_set ()) {}
        if (!ready.load ())
        {
            db.fallback = build_fallback ();
            ready.store (true);
        }
        lock.clear ();
    }
    return &db;
}

ChafaTermInfo *
chafa_term_db_get_fallback_info (ChafaTermDb *term_db)
{
    RETURN_VAL_IF_FAIL (term_db != NULL, NULL);
    return chafa_term_info_copy (term_db->fallback);
}

} /* extern "C" */

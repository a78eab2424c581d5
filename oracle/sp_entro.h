/* sp_entro.h -- the EntroCoder interface of the reference (src/EntroCoders.hx:8-24) as a C vtable.
 * TEST INFRASTRUCTURE ONLY (see oracle.h). */
#ifndef JSP_SP_ENTRO_H
#define JSP_SP_ENTRO_H
#include <stdint.h>

typedef struct entro entro;
struct entro {
    void (*destroy)(entro *);
    void (*preinit)(entro *);
    void (*renewI)(entro *);
    void (*decodeBegin)(entro *, const uint8_t *src, int len, int pos0);
    int  (*decodeClr)(entro *, int cxi);
    int  (*decodeN)(entro *, int ptype);
    int  (*decodeP)(entro *, int ptype);
    int  (*decodeX)(entro *);
    int  (*decodeBT)(entro *);
    int  (*decodeBN)(entro *);
    int  (*decodeSXY)(entro *, int n);
    int  (*decodeMX)(entro *);
    int  (*decodeMY)(entro *);
    int  (*canDecodeBool)(entro *);
    int  (*decodeBool)(entro *);
    int  (*differentConstantsFor16bpp)(entro *);
    /* not in the reference: set when the coder was asked for a symbol after it had already read past the
     * end of the data, or when a symbol search ran off its table (both impossible on a valid stream) */
    int  (*failed)(entro *);
    /* not in the reference: the frame loop reports a failure of its own (colour-context index out of range, run budget
     * spent).  Defined behaviour: the range coder decodes nothing more after a frame's first failure of any kind (each
     * later call returns 0 and leaves the models alone -- its arithmetic has no defined meaning from there); the rANS
     * coder, whose 32-bit state stays well defined on garbage, goes on until the frame loop's next check. */
    void (*fail)(entro *);
};

entro *entro_rc_new(void);            /* EntroCoderRC, EntroCoders.hx:31-180 */
entro *entro_ans_new(int f0val);      /* EntroCoderANS, EntroCoders.hx:182-313 */

enum { SP_MSR_X = 256, SP_MSR_Y = 256 };   /* ScreenPressor.hx:21-22 */
enum { CC_CXMAX = 4096, CC_NCXMAX = 6 };   /* EntroCoders.hx:26-29 */

#endif

/* pybpl_parsefile.h -- the packed on-disk / wire format for parses and bit-packed image sets
 * (written and read by pybpl_b200/packed.py).
 *
 * The reference keeps parses as Python object graphs (CharacterToken -> StrokeToken / RelationToken,
 * pybpl/objects/concept.py:217-241, part.py:188-249, relation.py:14-32; types concept.py:66-107) and loads its
 * library from MAT files (pybpl/library/library.py:169-178).  A parse file is one flat little-endian buffer whose
 * sections are, byte for byte, the device arrays of bpl_token_batch / bpl_token_type (include/pybpl_b200.h) and
 * image sets in the bit layout bpl_pack_images_* produces, so that loading is one read into pinned memory and one
 * host->device copy.  Every section starts on a 16-byte boundary.
 *
 * Sections (name, element type, shape); P tokens, K strokes, S sub-strokes, T images:
 *   required   tok_stroke_off i32[P+1]  stroke_sub_off i32[K+1]  ctrl f32[S][5][2]  invscale f32[S]
 *              position f32[K][2]  eps f32[P]  sigma f32[P]
 *   optional   affine f32[P][4]
 *   optional   ctrl_type f32[S][5][2]  invscale_type f32[S]  rel_cat i32[K]  attach_ix i32[K]  attach_subix i32[K]
 *              gpos f32[K][2]  eval_spot_type f32[K]  eval_spot f32[K]          (all eight or none)
 *   optional   image_bits u32[T][345]   pixel i of an image = bit (i & 31) of word (i >> 5), row-major
 */
#ifndef PYBPL_PARSEFILE_H
#define PYBPL_PARSEFILE_H
#include <stdint.h>

#define BPL_PARSEFILE_MAGIC "BPLPARS1"

enum { BPL_PF_F32 = 0, BPL_PF_I32 = 1, BPL_PF_U32 = 2 };

#pragma pack(push, 1)
typedef struct {
    char magic[8];          /* BPL_PARSEFILE_MAGIC */
    uint32_t header_bytes;  /* offset of the first section (multiple of 16) */
    uint32_t n_sections;
    int32_t P, K, S, T;     /* tokens, strokes, sub-strokes, images */
    int32_t ncpt, neval;    /* 5, points per sub-stroke (200) */
    int32_t reserved[2];
} bpl_parsefile_header;     /* 48 bytes, followed by n_sections entries */

typedef struct {
    char name[16];          /* zero padded */
    uint32_t dtype;         /* BPL_PF_* */
    uint32_t ndim;          /* 1..3 */
    uint32_t shape[3];      /* unused dimensions 0 */
    uint64_t offset;        /* from the start of the file, multiple of 16 */
    uint64_t nbytes;
} bpl_parsefile_section;    /* 52 bytes */
#pragma pack(pop)

#endif

// Declaration-only stand-in for <png.h> -- TEST INFRASTRUCTURE ONLY.
//
// libpng's headers are not installed in this image.  ndb::Buffer's PNG members
// (lib/gpc/buffer.hpp:197-474, :772-875) are member functions of a class template and are
// never instantiated by oracle/ref_harness.cpp, so declarations are all the reference
// headers need; nothing here is ever called or linked.
#ifndef GPC_ORACLE_PNG_SHIM
#define GPC_ORACLE_PNG_SHIM
#include <csetjmp>
#include <cstddef>
#include <cstdio>

typedef unsigned char png_byte;
typedef png_byte* png_bytep;
typedef png_byte** png_bytepp;
typedef struct png_struct_def png_struct;
typedef png_struct* png_structp;
typedef struct png_info_def png_info;
typedef png_info* png_infop;
typedef unsigned int png_uint_32;
typedef void* png_voidp;
typedef void (*png_error_ptr)(png_structp, const char*);

#define PNG_LIBPNG_VER_STRING "shim"
#define PNG_COLOR_TYPE_GRAY 0
#define PNG_COLOR_TYPE_RGB 2
#define PNG_COLOR_TYPE_RGBA 6
#define PNG_INTERLACE_NONE 0
#define PNG_COMPRESSION_TYPE_BASE 0
#define PNG_FILTER_TYPE_BASE 0

extern "C" {
int png_sig_cmp(png_bytep, size_t, size_t);
png_structp png_create_read_struct(const char*, png_voidp, png_error_ptr, png_error_ptr);
png_structp png_create_write_struct(const char*, png_voidp, png_error_ptr, png_error_ptr);
png_infop png_create_info_struct(png_structp);
jmp_buf* png_set_longjmp_fn(png_structp, void (*)(jmp_buf, int), size_t);
void png_init_io(png_structp, FILE*);
void png_set_sig_bytes(png_structp, int);
void png_read_info(png_structp, png_infop);
png_uint_32 png_get_image_width(png_structp, png_infop);
png_uint_32 png_get_image_height(png_structp, png_infop);
png_byte png_get_color_type(png_structp, png_infop);
png_byte png_get_bit_depth(png_structp, png_infop);
int png_set_interlace_handling(png_structp);
void png_read_update_info(png_structp, png_infop);
size_t png_get_rowbytes(png_structp, png_infop);
void png_read_image(png_structp, png_bytepp);
void png_set_IHDR(png_structp, png_infop, png_uint_32, png_uint_32, int, int, int, int, int);
void png_write_info(png_structp, png_infop);
void png_write_image(png_structp, png_bytepp);
void png_write_end(png_structp, png_infop);
}
#define png_jmpbuf(p) (*png_set_longjmp_fn((p), longjmp, sizeof(jmp_buf)))
#endif

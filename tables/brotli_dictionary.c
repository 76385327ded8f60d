/* RFC 7932 Appendix A static dictionary (122 784 bytes), embedded from the committed
 * binary tables/brotli_dictionary.bin (generated + CRC-checked by tables/gen_tables.py).
 * Reference counterpart: kBrotliDictionary, src/dictionary/mod.rs:18.
 * Compile with -DBROTLI_DICT_PATH="\"/abs/path/brotli_dictionary.bin\"". */
#include <stdint.h>
#ifndef BROTLI_DICT_PATH
#error "BROTLI_DICT_PATH must point at tables/brotli_dictionary.bin"
#endif
__asm__(".section .rodata\n"
        ".global kBrotliDictionaryData\n"
        ".balign 16\n"
        "kBrotliDictionaryData:\n"
        ".incbin \"" BROTLI_DICT_PATH "\"\n"
        ".global kBrotliDictionaryDataEnd\n"
        "kBrotliDictionaryDataEnd:\n"
        ".byte 0\n"
        ".previous\n");

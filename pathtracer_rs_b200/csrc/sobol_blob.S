/* Links pathtracer_rs_b200/data/sobol_tables.bin into libptrs_b200.so (path given by -DSOBOL_BLOB_PATH). */
    .section .rodata
    .balign 16
    .global ptrs_sobol_blob
    .global ptrs_sobol_blob_end
ptrs_sobol_blob:
    .incbin SOBOL_BLOB_PATH
ptrs_sobol_blob_end:
    .byte 0
    .section .note.GNU-stack,"",@progbits

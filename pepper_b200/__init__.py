"""pepper_b200 — B200-native (sm_100a) drop-in for the hot path of kishwarshafin/pepper: pileup summary encoders + recurrent
network inference, and the stages either side of it, behind the C-ABI of include/pepper_b200.h (libpepper_b200.so).

    bamio      host BGZF/BAM/BAI + FASTA/FAI readers            (pb_bam_*, pb_fasta_*)
    reads      batched BAM_handler.get_reads on the GPU          (pb_get_reads_*)
    realign    SSW-exact read -> reference realignment           (pb_realign_*)
    variant    RegionalSummaryGenerator encoder, bi-LSTM + MLP   (pb_variant_*)
    polish     SummaryGenerator encoder, chunking, bi-GRU, stitch (pb_polish_*)
    pipeline   fused make_images + inference callers
    frontend   files -> predictions / candidate records / polished sequence in one call
    candidates find_candidates' per-record selection             (pb_variant_find_candidates_*)
    datastore  the reference's HDF5 store layouts
    dist       region sharding + prediction all-gather (torch.distributed)
    build      PEPPER / PEPPER_VARIANT: the reference's extension-module class names on top of the above
    synth, synth_files, weights   seeded synthetic inputs / files / weights for tests and benchmarks

There is no CPU fallback: the compute entry points raise without the CUDA library or a GPU (see _lib.py)."""

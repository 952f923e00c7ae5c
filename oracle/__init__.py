"""TEST INFRASTRUCTURE ONLY. See oracle/README.md. Not imported by image_compression_2_b200."""

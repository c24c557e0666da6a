"""GATK-emulation helpers on the hot path (the reference's kbbq/gatk is a namespace package)."""

#!/usr/bin/env python
"""Golden output of the REFERENCE's scripts/FastqToSam.pl (the command runRufus.sh:159-161 writes into the `.generator`
of a FASTQ input) on a hand-made FASTQ with the cases that matter for a byte-exact drop-in: names with spaces and tabs,
leading whitespace, CRLF line ends, an empty name line, a record cut short by the end of the file, no trailing newline.
Run in the build container (needs perl and /root/reference); writes tests/golden/fastq_to_sam/{edge.fastq,edge.sam}."""
import os, subprocess
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "fastq_to_sam")
os.makedirs(OUT, exist_ok=True)
fq = (b"@read1/1 extra words\nACGTN\n+\nIIIII\n"
      b"@read2\tafter a tab\nacgt\n+read2\n!!!!\n"
      b"  @lead/2\nAC\n+\nII\n"
      b"@crlf/1 x\r\nACGT\r\n+\r\nIIII\r\n"
      b"\nGG\n+\nII\n"
      b"@\nA\n+\nI\n"
      b"@x\n\n+\n\n"
      b"@onlyname\r\nTT\n+\nII\n"
      b"@cut/1\nACGTACGT\n+")
open(os.path.join(OUT, "edge.fastq"), "wb").write(fq)
sam = subprocess.run(["perl", "/root/reference/scripts/FastqToSam.pl", os.path.join(OUT, "edge.fastq")], check=True, capture_output=True).stdout
open(os.path.join(OUT, "edge.sam"), "wb").write(sam)
print(sam.decode(errors="replace"))

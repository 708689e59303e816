"""`cfsan_snp_pipeline`-compatible dispatcher for the hot-path subcommands.

Mirror of snppipeline/cfsan_snp_pipeline.py for filter_regions / merge_sites / call_consensus / snp_matrix / distance /
snp_reference: the same flags, defaults and validators (cfsan_snp_pipeline.py:307-457, :528-543), the same `func` /
`excepthook` binding (:323-324, :339-340, :409-410, :442-443, :456-457) and the same entry points (parse_argument_list, run_command_from_args, run_command_from_arg_list,
run_command_from_line, main).  The other subcommands are outside this build.

    python -m snp_pipeline_b200.cfsan_snp_pipeline call_consensus -l snplist.txt -o s1/consensus.fasta s1/reads.all.pileup
"""
from __future__ import annotations

import argparse
import sys

from . import __version__
from . import call_consensus, distance, filter_regions, merge_sites, snp_matrix, snp_reference, utils


def _min_cons_freq(value):
    fvalue = float(value)
    if fvalue <= 0.5 or fvalue > 1:
        raise argparse.ArgumentTypeError("Minimum consensus frequency must be > 0.5 and <= 1.0")
    return fvalue


def _min_cons_strand_bias(value):
    fvalue = float(value)
    if fvalue < 0.0 or fvalue > 0.5:
        raise argparse.ArgumentTypeError("Minimum consensus strand bias must be >= 0.0 and <= 0.5")
    return fvalue


def parse_argument_list(argv):
    """argv without the program name -> Namespace (subparser_name, func, excepthook + the step's fields)."""
    fc = argparse.ArgumentDefaultsHelpFormatter
    parser = argparse.ArgumentParser(prog="cfsan_snp_pipeline",
                                     description="The CFSAN SNP Pipeline's pileup -> consensus -> SNP-matrix -> "
                                                 "distance steps on an NVIDIA B200 (snp_pipeline_b200).")
    parser.add_argument("--version", action="version", version="%(prog)s version " + __version__)
    subparsers = parser.add_subparsers(dest="subparser_name", help=None, metavar="subcommand")
    subparsers.required = True
    ver = dict(action="version", version="%(prog)s version " + __version__)

    sp = subparsers.add_parser("filter_regions", help="Remove abnormally dense SNPs from all samples", formatter_class=fc,
                               description="Remove abnormally dense SNPs from the input VCF file, save the reserved SNPs into "
                                           "a new VCF file, and save the removed SNPs into another VCF file.",
                               epilog='You can filter snps more than once by specifying multiple window sizes and max snps.  '
                                      'For example "-m 3 2 -w 1000 100" will filter more than 3 snps in 1000 bases and also '
                                      'more than 2 snps in 100 bases.')
    sp.add_argument(dest="sampleDirsFile", type=str, help="File containing a list of directories -- one per sample")
    sp.add_argument(dest="refFastaFile", type=str, help="Relative or absolute path to the reference fasta file")
    sp.add_argument("-f", "--force", dest="forceFlag", action="store_true",
                    help="Force processing even when result files already exist and are newer than inputs")
    sp.add_argument("-n", "--vcfname", dest="vcfFileName", type=str, default="var.flt.vcf", metavar="NAME",
                    help="File name of the input VCF files which must exist in each of the sample directories")
    sp.add_argument("-l", "--edge_length", dest="edgeLength", type=int, default=500, metavar="EDGE_LENGTH",
                    help="The length of the edge regions in a contig, in which all SNPs will be removed.")
    sp.add_argument("-w", "--window_size", dest="windowSizeList", type=int, default=[1000], nargs="*", metavar="WINDOW_SIZE",
                    help="The length of the window in which the number of SNPs should be no more than max_num_snp.")
    sp.add_argument("-m", "--max_snp", dest="maxSnpsList", type=int, default=[3], nargs="*", metavar="MAX_NUM_SNPs",
                    help="The maximum number of SNPs allowed in a window.")
    sp.add_argument("-g", "--out_group", dest="outGroupFile", type=str, default=None, metavar="OUT_GROUP",
                    help="Relative or absolute path to the file indicating outgroup samples, one sample ID per line.")
    sp.add_argument("-M", "--mode", dest="mode", choices=["all", "each"], default="all",
                    help="Control whether dense snp regions found in any sample are filtered from all of the samples, or "
                         "each sample independently.")
    sp.add_argument("-v", "--verbose", dest="verbose", type=int, default=1, metavar="0..5",
                    help="Verbose message level (0=no info, 5=lots)")
    sp.add_argument("--version", **ver)
    sp.set_defaults(func=filter_regions.filter_regions, excepthook=utils.handle_global_exception)

    sp = subparsers.add_parser("merge_sites", help="Prepare the list of sites having SNPs", formatter_class=fc,
                               description="Combine the SNP positions across all samples into a single unified SNP "
                                           "list file identifying the positions and sample names where SNPs were called.")
    sp.add_argument(dest="sampleDirsFile", type=str, help="File containing a list of directories -- one per sample")
    sp.add_argument(dest="filteredSampleDirsFile", type=str,
                    help="Output file that will be created containing the filtered list of sample directories")
    sp.add_argument("-f", "--force", dest="forceFlag", action="store_true",
                    help="Force processing even when result file already exists and is newer than inputs")
    sp.add_argument("-n", "--vcfname", dest="vcfFileName", type=str, default="var.flt.vcf", metavar="NAME",
                    help="File name of the VCF files which must exist in each of the sample directories")
    sp.add_argument("--maxsnps", dest="maxSnps", type=int, default=-1, metavar="INT",
                    help="Exclude samples having more than this maximum allowed number of SNPs. -1 disables.")
    sp.add_argument("-o", "--output", dest="snpListFile", type=str, default="snplist.txt", metavar="FILE",
                    help="Output file.  Relative or absolute path to the SNP list file")
    sp.add_argument("-v", "--verbose", dest="verbose", type=int, default=1, metavar="0..5",
                    help="Verbose message level (0=no info, 5=lots)")
    sp.add_argument("--version", **ver)
    sp.set_defaults(func=merge_sites.merge_sites, excepthook=utils.handle_global_exception)

    sp = subparsers.add_parser("call_consensus", help="Call the consensus base at high-confidence sites",
                               formatter_class=fc,
                               description="Call the consensus base for a sample at the specified positions where "
                                           "high-confidence SNPs were previously called in any of the samples.")
    sp.add_argument(dest="allPileupFile", type=str, help="Path to the genome-wide pileup file for this sample.")
    sp.add_argument("-f", "--force", dest="forceFlag", action="store_true",
                    help="Force processing even when result file already exists and is newer than inputs.")
    sp.add_argument("-l", "--snpListFile", dest="snpListFile", type=str, default="snplist.txt", metavar="FILE",
                    help="Path to the SNP list file across all samples.")
    sp.add_argument("-e", "--excludeFile", dest="excludeFile", type=str, default=None, metavar="FILE",
                    help="VCF file of positions to exclude.")
    sp.add_argument("-o", "--output", dest="consensusFile", type=str, default="consensus.fasta", metavar="FILE",
                    help="Output file. Path to the consensus fasta file for this sample.")
    sp.add_argument("-q", "--minBaseQual", dest="minBaseQual", type=int, default=0, metavar="INT",
                    help="Mimimum base quality score to count a read.")
    sp.add_argument("-c", "--minConsFreq", dest="minConsFreq", type=_min_cons_freq, default=0.60, metavar="FREQ",
                    help="Consensus frequency. Mimimum fraction of high-quality reads supporting the consensus.")
    sp.add_argument("-D", "--minConsDpth", dest="minConsDpth", type=int, default=1, metavar="INT",
                    help="Consensus depth. Minimum number of high-quality reads supporting the consensus.")
    sp.add_argument("-d", "--minConsStrdDpth", dest="minConsStrdDpth", type=int, default=0, metavar="INT",
                    help="Consensus strand depth, required on both strands.")
    sp.add_argument("-b", "--minConsStrdBias", dest="minConsStrdBias", type=_min_cons_strand_bias, default=0,
                    metavar="FREQ", help="Strand bias. Minimum fraction of consensus-supporting reads on each strand.")
    sp.add_argument("--vcfFileName", dest="vcfFileName", type=str, default=None, metavar="NAME",
                    help="VCF Output file name.  If not set, no VCF file is written.")
    sp.add_argument("--vcfRefName", dest="vcfRefName", type=str, default="Unknown reference", metavar="NAME",
                    help="Name of the reference file.  Only used in the generated VCF file header.")
    sp.add_argument("--vcfAllPos", dest="vcfAllPos", action="store_true",
                    help="Parse every pileup position, not just the snp positions.")
    sp.add_argument("--vcfPreserveRefCase", dest="vcfPreserveRefCase", action="store_true",
                    help="Emit each reference base in the case it has in the reference.")
    sp.add_argument("--vcfFailedSnpGt", dest="vcfFailedSnpGt", type=str, default=".", choices=[".", "0", "1"],
                    help="Controls the VCF file GT data element when a snp fails filters.")
    sp.add_argument("-v", "--verbose", dest="verbose", type=int, default=1, metavar="0..5",
                    help="Verbose message level (0=no info, 5=lots)")
    sp.add_argument("--version", **ver)
    sp.set_defaults(func=call_consensus.call_consensus, excepthook=utils.handle_sample_exception)

    sp = subparsers.add_parser("snp_matrix", help="Create a matrix of SNPs", formatter_class=fc,
                               description="Create the SNP matrix containing the consensus base for each of the "
                                           "samples at the positions where high-confidence SNPs were found.")
    sp.add_argument(dest="sampleDirsFile", type=str, help="File containing a list of directories -- one per sample")
    sp.add_argument("-f", "--force", dest="forceFlag", action="store_true",
                    help="Force processing even when result file already exists and is newer than inputs")
    sp.add_argument("-c", "--consFileName", dest="consFileName", type=str, default="consensus.fasta", metavar="NAME",
                    help="File name of the consensus SNP call file which must exist in each of the sample directories")
    sp.add_argument("-o", "--output", dest="snpmaFile", type=str, default="snpma.fasta", metavar="FILE",
                    help="Output file.  Relative or absolute path to the SNP matrix file")
    sp.add_argument("-v", "--verbose", dest="verbose", type=int, default=1, metavar="0..5",
                    help="Verbose message level (0=no info, 5=lots)")
    sp.add_argument("--version", **ver)
    sp.set_defaults(func=snp_matrix.create_snp_matrix, excepthook=utils.handle_global_exception)

    sp = subparsers.add_parser("distance", help="Calculate the SNP distances between samples", formatter_class=fc,
                               description="Calculate pairwise SNP distances from the multi-fasta SNP matrix.")
    sp.add_argument(dest="inputFile", type=str, metavar="snpMatrixFile",
                    help="Relative or absolute path to the input multi-fasta SNP matrix file.")
    sp.add_argument("-f", "--force", dest="forceFlag", action="store_true",
                    help="Force processing even when result file already exists and is newer than inputs")
    sp.add_argument("-p", "--pairs", dest="pairwiseFile", type=str, default=None, metavar="FILE",
                    help="Relative or absolute path to the pairwise distance output file.")
    sp.add_argument("-m", "--matrix", dest="matrixFile", type=str, default=None, metavar="FILE",
                    help="Relative or absolute path to the distance matrix output file.")
    sp.add_argument("-v", "--verbose", dest="verbose", type=int, default=1, metavar="0..5",
                    help="Verbose message level (0=no info, 5=lots)")
    sp.add_argument("--version", **ver)
    sp.set_defaults(func=distance.calculate_snp_distances, excepthook=utils.handle_global_exception)

    sp = subparsers.add_parser("snp_reference", help="Write reference bases at SNP locations to a fasta file",
                               formatter_class=fc,
                               description="Write reference sequence bases at SNP locations to a fasta file.")
    sp.add_argument(dest="referenceFile", type=str,
                    help="Relative or absolute path to the reference bases file in fasta format")
    sp.add_argument("-f", "--force", dest="forceFlag", action="store_true",
                    help="Force processing even when result file already exists and is newer than inputs")
    sp.add_argument("-l", "--snpListFile", dest="snpListFile", type=str, default="snplist.txt", metavar="FILE",
                    help="Relative or absolute path to the SNP list file")
    sp.add_argument("-o", "--output", dest="snpRefFile", type=str, default="referenceSNP.fasta", metavar="FILE",
                    help="Output file.  Relative or absolute path to the SNP reference sequence file")
    sp.add_argument("-v", "--verbose", dest="verbose", type=int, default=1, metavar="0..5",
                    help="Verbose message level (0=no info, 5=lots)")
    sp.add_argument("--version", **ver)
    sp.set_defaults(func=snp_reference.create_snp_reference_seq, excepthook=utils.handle_global_exception)

    args = parser.parse_args(argv)

    if args.subparser_name == "filter_regions":             # cfsan_snp_pipeline.py:529-543
        if len(args.windowSizeList) != len(args.maxSnpsList):
            utils.global_error("Error: you must specify the same number of arguments for window size and max snps.")
        for window_size in args.windowSizeList:
            if window_size < 1:
                utils.global_error("Error: the length of the window must be a positive integer, and the input is %d." % window_size)
        for max_snps in args.maxSnpsList:
            if max_snps < 1:
                utils.global_error("Error: the maximum number of SNPs allowed must be a positive integer, and the input is %d." % max_snps)
        if args.edgeLength < 1:
            utils.global_error("Error: the length of the edge regions must be a positive integer, and the input is %d." % args.edgeLength)
    return args


def parse_command_line(line):
    """Command line without the program name, as one string (what the reference's unit tests use)."""
    return parse_argument_list(line.split())


def run_command_from_args(args):
    """cfsan_snp_pipeline.py:568-589: install the step's excepthook, set verbosity, run the step."""
    sys.excepthook = args.excepthook
    utils.set_logging_verbosity(args)
    args.func(args)


def run_command_from_arg_list(argv):
    run_command_from_args(parse_argument_list(argv))


def run_command_from_line(line):
    run_command_from_arg_list(line.split())


def main():
    run_command_from_arg_list(sys.argv[1:])


if __name__ == "__main__":
    main()

"""Host-side helpers of the hot path: the reference's conventions for errors, verbosity, freshness and the small
file formats either side of the kernels.  Mirrors the parts of snppipeline/utils.py that the four hot-path steps
use (same function names, arguments and behaviour); nothing here computes on the hot path.

  global_error / sample_error / handle_*_exception   utils.py:542-726   exit 100 / 98 protocol, $errorOutputFile
  verify_existing_input_files / verify_non_empty_input_files   utils.py:868-925
  target_needs_rebuild                                 utils.py:977-1009 make-style mtime freshness
  write_list_of_snps / read_snp_position_list          utils.py:1056-1088
  convert_vcf_file_to_snp_set                          utils.py:1113-1132 (PyVCF3's Reader restated: CHROM, POS only)
"""
from __future__ import annotations

import os
import re
import sys
import time
import traceback

log_verbosity = 0


def set_logging_verbosity(args):
    """utils.py:39-53: pick up args.verbose."""
    global log_verbosity
    log_verbosity = getattr(args, "verbose", 0) or 0


def verbose_print(*args):
    if log_verbosity > 0:
        print(*args)


def timestamp():
    return time.strftime("%Y-%m-%d %H:%M:%S")


def program_name():
    return os.path.basename(sys.argv[0]) if sys.argv and sys.argv[0] else "cfsan_snp_pipeline"


def program_name_with_command():
    """utils.py:147-170: `cfsan_snp_pipeline <subcommand>` when run through the dispatcher."""
    name = program_name()
    if name == "cfsan_snp_pipeline" and len(sys.argv) > 1:
        name += " " + sys.argv[1]
    return name


def command_line_short():
    return " ".join([program_name()] + sys.argv[1:])


def print_log_header(classpath=False):
    """utils.py:85-127, minus the locale.format call that no longer exists on Python >= 3.12."""
    verbose_print("# Command           : %s" % " ".join(sys.argv))
    verbose_print("# Working Directory : %s" % os.getcwd())
    verbose_print("# Hostname          : %s" % os.uname().nodename)
    verbose_print("# Program Version   : %s %s (snp_pipeline_b200)" % (program_name(), _version()))
    verbose_print("# %s" % timestamp())
    verbose_print("")


def _version():
    from . import __version__
    return __version__


def print_arguments(args):
    """utils.py:130-144."""
    verbose_print("Options:")
    options_dict = vars(args)
    for key in sorted(options_dict):
        if key in ("subparser_name", "func", "excepthook"):
            continue
        verbose_print("    %s=%s" % (key, options_dict[key]))
    verbose_print("")


# ----------------------------------------------------------------------------- error protocol (utils.py:542-726)
def _err_log(lines):
    path = os.environ.get("errorOutputFile")
    if path:
        with open(path, "a") as err_log:
            for ln in lines:
                print(ln, file=err_log)
            print("=" * 80, file=err_log)


def report_error(message):
    _err_log(["%s failed." % program_name_with_command()] + ([message] if message else []))
    sys.stdout.flush()
    if message:
        print(message, file=sys.stderr)


def global_error(message):
    """Fatal for the whole run: log, exit 100 whatever StopOnSampleError says."""
    report_error(message)
    sys.exit(100)


def sample_error(message, continue_possible=False):
    """Per-sample error: exit 100 when StopOnSampleError is unset/true; otherwise exit 98 unless the step can go on."""
    stop_env = os.environ.get("StopOnSampleError")
    stop = stop_env is None or stop_env == "true"
    head = "%s failed." % program_name_with_command() if (stop or not continue_possible) else program_name_with_command()
    _err_log([head, message])
    sys.stdout.flush()
    print(message, file=sys.stderr)
    if stop:
        sys.exit(100)
    if not continue_possible:
        sys.exit(98)


def _log_exception(exc_type, exc_value, exc_traceback):
    entries = traceback.extract_tb(exc_traceback)
    lines = ["Error detected while running %s." % program_name_with_command(), "", "The command line was:",
             "    %s" % command_line_short(), ""]
    if entries:
        file_name, line_number, function_name, code_text = entries[-1]
        lines.append("%s exception in function %s at line %d in file %s" % (exc_type.__name__, function_name,
                                                                              line_number, file_name))
        lines.append("    %s" % code_text)
    _err_log(lines)
    sys.stdout.flush()
    traceback.print_exception(exc_type, exc_value, exc_traceback)


def handle_global_exception(exc_type, exc_value, exc_traceback):
    """Uncaught exception in merge_sites / snp_matrix / distance (CUDA errors included) -> exit 100."""
    _log_exception(exc_type, exc_value, exc_traceback)
    sys.exit(100)


def handle_sample_exception(exc_type, exc_value, exc_traceback):
    """Uncaught exception in call_consensus -> exit 100, or 98 when StopOnSampleError=false."""
    _log_exception(exc_type, exc_value, exc_traceback)
    stop_env = os.environ.get("StopOnSampleError")
    sys.exit(100 if (stop_env is None or stop_env == "true") else 98)


# ----------------------------------------------------------------------------- input checks (utils.py:868-925)
def verify_existing_input_files(error_prefix, file_list, error_handler=None, continue_possible=False):
    bad = 0
    for path in file_list:
        if not os.path.isfile(path):
            bad += 1
            _file_problem("%s %s does not exist." % (error_prefix, path), error_handler, continue_possible)
    return bad


def verify_non_empty_input_files(error_prefix, file_list, error_handler=None, continue_possible=False):
    bad = 0
    for path in file_list:
        if not os.path.isfile(path):
            bad += 1
            _file_problem("%s %s does not exist." % (error_prefix, path), error_handler, continue_possible)
        elif os.path.getsize(path) == 0:
            bad += 1
            _file_problem("%s %s is empty." % (error_prefix, path), error_handler, continue_possible)
    return bad


def _file_problem(message, error_handler, continue_possible):
    if error_handler == "global":
        global_error(message)
    elif error_handler == "sample":
        sample_error(message, continue_possible)
    else:
        print(message, file=sys.stderr)
        path = os.environ.get("errorOutputFile")
        if path:
            with open(path, "a") as err_log:
                print(message, file=err_log)


def target_needs_rebuild(source_files, target_file):
    """utils.py:977-1009: rebuild when the target is missing, empty, or older than any existing source."""
    if not os.path.isfile(target_file) or os.path.getsize(target_file) == 0:
        return True
    target_timestamp = os.stat(target_file).st_mtime
    for source in source_files:
        if os.path.isfile(source) and os.stat(source).st_mtime > target_timestamp:
            return True
    return False


# ----------------------------------------------------------------------------- small file formats
def write_list_of_snps(file_path, snp_dict):
    """utils.py:1056-1070: snplist.txt, lines sorted by (chrom string, position)."""
    with open(file_path, "w") as snp_list_file:
        for key in sorted(snp_dict):
            names = snp_dict[key]
            snp_list_file.write("%s\t%d\t%d\t%s\n" % (key[0], key[1], len(names), "\t".join(names)))


def read_snp_position_list(snp_list_file_path):
    """utils.py:1073-1088: [(chrom, pos)] in file order; a malformed line raises (-> exit 100 / 98)."""
    snp_list = []
    with open(snp_list_file_path, "r") as snp_list_file:
        for line in snp_list_file:
            chrom, pos = line.split()[0:2]
            snp_list.append((chrom, int(pos)))
    return snp_list


_VCF_ROW_SPLIT = re.compile("\t| +")


def read_vcf_positions(vcf_file_path):
    """[(CHROM, POS)] of every data line, as PyVCF3 1.0.3's vcf.Reader reports them (the reference consumes nothing
    else, utils.py:1127-1131): lines are stripped and blank ones dropped, `##` meta lines and the one header line
    that follows are skipped, data rows split on tab or runs of spaces and must have at least 8 columns."""
    out = []
    with open(vcf_file_path, "r") as f:
        lines = [ln for ln in (raw.strip() for raw in f) if ln]
    i = 0
    while i < len(lines) and lines[i].startswith("##"):
        i += 1
    if i >= len(lines):
        raise StopIteration("vcf file %s holds no header line" % vcf_file_path)
    for ln in lines[i + 1:]:
        row = _VCF_ROW_SPLIT.split(ln.rstrip())
        if len(row) < 8:
            raise IndexError("list index out of range")
        out.append((row[0], int(row[1])))
    return out


def convert_vcf_file_to_snp_set(vcf_file_path):
    """utils.py:1113-1132."""
    return set(read_vcf_positions(vcf_file_path))


def sample_id_from_file(file_path):
    """utils.py:459-485: the sample id is the name of the file's directory."""
    return os.path.basename(os.path.dirname(os.path.abspath(file_path)))


def fasta_contig_lengths(fasta_path):
    """{record id: len(record.seq)} as the SeqIO.parse loop of filter_regions.py:183-192 builds it (a repeated id keeps the
    last length, lines in front of the first '>' are skipped)."""
    lengths, cur = {}, None
    with open(fasta_path, "r") as f:
        for line in f:
            if line.startswith(">"):
                words = line[1:].split(None, 1)
                if not words:
                    raise IndexError("list index out of range")
                cur = words[0]
                lengths[cur] = 0
            elif cur is not None:
                lengths[cur] += len("".join(line.split()))
    return lengths


def fasta_record_text(seq_id, seq):
    """Bio.SeqIO's FastaWriter as call_consensus.py:189-192 drives it (description ""): 60-column lines."""
    parts = [">%s\n" % seq_id]
    for i in range(0, len(seq), 60):
        parts.append(seq[i:i + 60])
        parts.append("\n")
    return "".join(parts)

/* ref_shim: the libpcap entry points vtkPacketFileReader.h / vtkPacketFileWriter.cxx call, for
 * classic little-endian microsecond pcap files read and written through stdio, so that
 * fgetpos/fsetpos on pcap_file() behave as they do with libpcap's own savefile reader
 * (test infrastructure). */
#ifndef REF_SHIM_PCAP_H
#define REF_SHIM_PCAP_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#define PCAP_ERRBUF_SIZE 256
#define DLT_EN10MB 1
typedef unsigned char u_char;
typedef unsigned int bpf_u_int32;
struct pcap_pkthdr { struct timeval ts; bpf_u_int32 caplen; bpf_u_int32 len; };
struct bpf_program { int unused; };
typedef struct pcap {
  FILE* f;
  unsigned char* buf;
  size_t cap;
  struct pcap_pkthdr hdr;
  char err[PCAP_ERRBUF_SIZE];
  int linktype, snaplen;
} pcap_t;
typedef struct pcap_dumper { FILE* f; } pcap_dumper_t;

static inline pcap_t* pcap_open_offline(const char* name, char* errbuf) {
  FILE* f = fopen(name, "rb");
  unsigned char gh[24];
  pcap_t* p;
  if (!f) { snprintf(errbuf, PCAP_ERRBUF_SIZE, "%s: cannot open", name); return 0; }
  if (fread(gh, 1, 24, f) != 24 || !(gh[0] == 0xd4 && gh[1] == 0xc3 && gh[2] == 0xb2 && gh[3] == 0xa1)) {
    snprintf(errbuf, PCAP_ERRBUF_SIZE, "%s: not a pcap file", name); fclose(f); return 0;
  }
  p = (pcap_t*)calloc(1, sizeof(pcap_t));
  p->f = f; p->cap = 65536; p->buf = (unsigned char*)malloc(p->cap);
  return p;
}
static inline int pcap_compile(pcap_t*, struct bpf_program*, const char*, int, bpf_u_int32) { return 0; }
static inline int pcap_setfilter(pcap_t*, struct bpf_program*) { return 0; }
static inline char* pcap_geterr(pcap_t* p) { return p->err; }
static inline FILE* pcap_file(pcap_t* p) { return p->f; }
static inline void pcap_close(pcap_t* p) { if (p) { if (p->f) fclose(p->f); free(p->buf); free(p); } }
static inline int pcap_next_ex(pcap_t* p, struct pcap_pkthdr** h, const u_char** data) {
  unsigned int rh[4];
  if (fread(rh, 4, 4, p->f) != 4) return -2;
  if (rh[2] > p->cap) { p->cap = rh[2]; p->buf = (unsigned char*)realloc(p->buf, p->cap); }
  if (fread(p->buf, 1, rh[2], p->f) != rh[2]) return -2;
  p->hdr.ts.tv_sec = rh[0]; p->hdr.ts.tv_usec = rh[1]; p->hdr.caplen = rh[2]; p->hdr.len = rh[3];
  *h = &p->hdr; *data = p->buf;
  return 1;
}
static inline pcap_t* pcap_open_dead(int linktype, int snaplen) {
  pcap_t* p = (pcap_t*)calloc(1, sizeof(pcap_t));
  p->linktype = linktype; p->snaplen = snaplen;
  return p;
}
static inline pcap_dumper_t* pcap_dump_open(pcap_t* p, const char* name) {
  FILE* f = fopen(name, "wb");
  unsigned int gh[6];
  pcap_dumper_t* d;
  if (!f) { snprintf(p->err, PCAP_ERRBUF_SIZE, "%s: cannot create", name); return 0; }
  gh[0] = 0xa1b2c3d4u; gh[1] = (4u << 16) | 2u; gh[2] = 0; gh[3] = 0; gh[4] = (unsigned)p->snaplen; gh[5] = (unsigned)p->linktype;
  fwrite(gh, 4, 6, f);
  d = (pcap_dumper_t*)malloc(sizeof(pcap_dumper_t));
  d->f = f;
  return d;
}
static inline void pcap_dump(u_char* user, const struct pcap_pkthdr* h, const u_char* sp) {
  pcap_dumper_t* d = (pcap_dumper_t*)user;
  unsigned int rh[4];
  rh[0] = (unsigned)h->ts.tv_sec; rh[1] = (unsigned)h->ts.tv_usec; rh[2] = h->caplen; rh[3] = h->len;
  fwrite(rh, 4, 4, d->f);
  fwrite(sp, 1, h->caplen, d->f);
}
static inline void pcap_dump_close(pcap_dumper_t* d) { if (d) { fclose(d->f); free(d); } }
#endif

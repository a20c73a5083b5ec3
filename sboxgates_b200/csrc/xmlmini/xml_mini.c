/* xmlmini/xml_mini.c -- a libxml2-free XML reader for the drop-in build.
 *
 * The reference needs libxml2 only to READ saved graphs (load_state, state.c:260-411, behind
 * --graph, -c and -d); it writes them with fprintf.  This file implements the five libxml2 entry
 * points state.c uses -- xmlParseFile, xmlGetProp, xmlFreeDoc, xmlFree, the node fields name /
 * children / next -- for files following gates.xsd: an optional <?xml ...?> declaration,
 * comments, nested elements with double- or single-quoted attributes; text content is ignored.
 * Linked into the GPU drop-in CLI (and into the reference build the oracle uses) so that loading
 * a partial graph, DOT and C output work where libxml2 is not installed.  The Python-side
 * counterpart with the same checks is sboxgates_b200/graph.py.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "libxml/parser.h"

static char *dup_range(const char *b, const char *e) {
  size_t n = (size_t)(e - b);
  char *s = malloc(n + 1);
  if (s != NULL) {
    memcpy(s, b, n);
    s[n] = '\0';
  }
  return s;
}

static void free_node(xmlNode *n) {
  while (n != NULL) {
    xmlNode *next = n->next;
    free_node(n->children);
    for (xmlAttrMini *a = n->attrs; a != NULL;) {
      xmlAttrMini *an = a->next;
      free(a->name);
      free(a->value);
      free(a);
      a = an;
    }
    free((void *)n->name);
    free(n);
    n = next;
  }
}

void xmlFreeDoc(xmlDocPtr doc) {
  if (doc == NULL) return;
  free_node(doc->children);
  free(doc);
}

void sbg_xml_free(void *p) { free(p); }

xmlChar *xmlGetProp(const xmlNode *node, const xmlChar *name) {
  for (const xmlAttrMini *a = node->attrs; a != NULL; a = a->next) {
    if (strcmp(a->name, (const char *)name) == 0) {
      return (xmlChar *)dup_range(a->value, a->value + strlen(a->value));
    }
  }
  return NULL;
}

static int is_name_char(int c) { return isalnum(c) || c == '_' || c == '-' || c == ':' || c == '.'; }

xmlDocPtr xmlParseFile(const char *filename) {
  FILE *fp = fopen(filename, "rb");
  if (fp == NULL) return NULL;
  fseek(fp, 0, SEEK_END);
  long len = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (len < 0 || len > (64L << 20)) { fclose(fp); return NULL; }
  char *buf = malloc((size_t)len + 1);
  if (buf == NULL) { fclose(fp); return NULL; }
  if (fread(buf, 1, (size_t)len, fp) != (size_t)len) { free(buf); fclose(fp); return NULL; }
  fclose(fp);
  buf[len] = '\0';

  xmlDoc *doc = calloc(1, sizeof(xmlDoc));
  xmlNode *stack[64];
  int depth = 0;
  xmlNode *last_top = NULL;
  int ok = (doc != NULL);
  const char *p = buf;
  while (ok && *p != '\0') {
    if (*p != '<') { p++; continue; }              /* text content: ignored */
    if (strncmp(p, "<?", 2) == 0) {                /* declaration / PI */
      const char *e = strstr(p, "?>");
      if (e == NULL) { ok = 0; break; }
      p = e + 2;
      continue;
    }
    if (strncmp(p, "<!--", 4) == 0) {
      const char *e = strstr(p, "-->");
      if (e == NULL) { ok = 0; break; }
      p = e + 3;
      continue;
    }
    if (p[1] == '/') {                             /* closing tag */
      const char *b = p + 2;
      const char *e = b;
      while (is_name_char((unsigned char)*e)) e++;
      if (depth == 0 || strlen((const char *)stack[depth - 1]->name) != (size_t)(e - b)
          || strncmp((const char *)stack[depth - 1]->name, b, (size_t)(e - b)) != 0) {
        ok = 0;
        break;
      }
      while (isspace((unsigned char)*e)) e++;
      if (*e != '>') { ok = 0; break; }
      depth--;
      p = e + 1;
      continue;
    }
    /* opening tag */
    const char *b = p + 1;
    const char *e = b;
    while (is_name_char((unsigned char)*e)) e++;
    if (e == b) { ok = 0; break; }
    xmlNode *node = calloc(1, sizeof(xmlNode));
    if (node == NULL) { ok = 0; break; }
    node->name = (const xmlChar *)dup_range(b, e);
    if (depth == 0) {
      if (last_top == NULL) doc->children = node; else last_top->next = node;
      last_top = node;
    } else {
      xmlNode *par = stack[depth - 1];
      if (par->last_child == NULL) par->children = node; else par->last_child->next = node;
      par->last_child = node;
    }
    xmlAttrMini *last_attr = NULL;
    int self_closing = 0;
    p = e;
    for (;;) {
      while (isspace((unsigned char)*p)) p++;
      if (*p == '/' && p[1] == '>') { self_closing = 1; p += 2; break; }
      if (*p == '>') { p++; break; }
      const char *nb = p;
      while (is_name_char((unsigned char)*p)) p++;
      if (p == nb) { ok = 0; break; }
      const char *ne = p;
      while (isspace((unsigned char)*p)) p++;
      if (*p != '=') { ok = 0; break; }
      p++;
      while (isspace((unsigned char)*p)) p++;
      char q = *p;
      if (q != '"' && q != '\'') { ok = 0; break; }
      const char *vb = ++p;
      while (*p != '\0' && *p != q) p++;
      if (*p != q) { ok = 0; break; }
      xmlAttrMini *a = calloc(1, sizeof(xmlAttrMini));
      if (a == NULL) { ok = 0; break; }
      a->name = dup_range(nb, ne);
      a->value = dup_range(vb, p);
      if (last_attr == NULL) node->attrs = a; else last_attr->next = a;
      last_attr = a;
      p++;
    }
    if (!ok) break;
    if (!self_closing) {
      if (depth >= 64) { ok = 0; break; }
      stack[depth++] = node;
    }
  }
  if (ok && depth != 0) ok = 0;
  free(buf);
  if (!ok) {
    xmlFreeDoc(doc);
    return NULL;
  }
  return doc;
}
